"""TEST INFRASTRUCTURE -- import-time stand-in for dgllife 0.2.8 (``requirements.txt:8``)."""
