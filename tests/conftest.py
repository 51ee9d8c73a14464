"""pytest configuration: markers and import paths.

The product package keeps the reference's flat module names (kmer, records,
data_file, constants, main) inside a directory whose name is not a Python
identifier, so the directory itself is put on sys.path.
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG_DIR = os.path.join(ROOT, "bioinformatics-project-for-shotgun-metagenomics-pseudo-alignment-shotgun-_b200")
for p in (ROOT, PKG_DIR, os.path.dirname(os.path.abspath(__file__))):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
