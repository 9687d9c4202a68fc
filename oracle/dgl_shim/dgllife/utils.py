"""TEST INFRASTRUCTURE -- ``one_hot_encoding`` as imported by ``src/utils/protein_featurizers.py:5``."""


def one_hot_encoding(x, allowable_set, encode_unknown=False):
    enc = [x == s for s in allowable_set]
    if encode_unknown:
        enc.append(x not in allowable_set)
    return enc
