"""
constants.py -- alphabets, legal quality characters and CLI defaults.

Same names and values as the reference's constants module
(/root/reference/src/constants.py:1-15); only the values matter to the path:
'N' is the null nucleotide (kmer.py:145), quality characters are ASCII 33..126,
defaults m = 1, p = 1, similarity threshold 0.95 (main.py:337-342).
"""
NULL_NUCLEOTIDES = {"N"}
NULL_NUCLEOTIDES_CHAR = "N"
REAL_NUCLEOTIDES = {"A", "C", "G", "T"}
REAL_NUCLEOTIDES_CHARS = "ACGT"
NUCLEOTIDES = REAL_NUCLEOTIDES | NULL_NUCLEOTIDES
NUCLEOTIDES_CHARS = "ACGTN"

# every printable, non-space ASCII character is a legal quality symbol; the score is the raw code point
PHRED33_SCORES = {chr(c): c for c in range(33, 127)}

DEFAULT_UNIQUE_THRESHOLD = 1
DEFAULT_AMBIGUOUS_THRESHOLD = 1
DEFAULT_SIMILARITY_THRESHOLD = 0.95
