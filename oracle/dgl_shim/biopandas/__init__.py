"""TEST INFRASTRUCTURE -- import-time stand-in for biopandas 0.2.8 (``requirements.txt:6``)."""
