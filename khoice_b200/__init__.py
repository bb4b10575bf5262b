"""khoice-b200: B200-native k-mer discriminatory-power path (khoice experiment type 1)."""
__version__ = "0.1.0"
