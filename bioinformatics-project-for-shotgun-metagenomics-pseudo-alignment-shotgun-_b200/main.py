#!/usr/bin/env python3
"""
main.py -- command line of the B200-native pseudo-alignment package.

Same tasks, flags, defaults, messages and exit behaviour as the reference CLI
(/root/reference/src/main.py:61-82, 317-402): `reference`, `dumpref`, `align`,
`dumpalign`; the misspelt `--ambiguous-threhold` is the real flag name and
`--reverse-complement` is accepted and ignored, as in the reference.  The work
itself is done by kmer.KmerReference / kmer.PseudoAlignment on the GPU.

One deliberate difference: `-t align -g G.fa -k K --reads R.fq -a OUT.aln` without
`-r` crashes in the reference with a TypeError (it calls save(None), main.py:372);
here the reference database is simply not written when no `-r` path is given.

Multi-GPU: launched under torchrun (`python -m torch.distributed.run --nproc-per-node N main.py ...`) every rank runs the
same task; KmerReference builds its index partitioned over the ranks and PseudoAlignment aligns one block of the reads
per rank (multi_gpu.init).  Rank 0 alone prints and writes files; the output is byte for byte the single-GPU output.
"""
import argparse
import gzip
import json
import os
import sys
from typing import List, Optional

from constants import DEFAULT_AMBIGUOUS_THRESHOLD, DEFAULT_SIMILARITY_THRESHOLD, DEFAULT_UNIQUE_THRESHOLD
from data_file import FASTAFile, FASTAQFile, InvalidExtensionError, NoRecordsInDataFile
from kmer import AddingExistingRead, KmerReference, NotValidatingUniqueMapping, PseudoAlignment

BAD_FORMAT = "Error: Incorrect format of input file."


def _is_root() -> bool:
    """Rank 0 of a torchrun launch (or the only process): the one that prints and writes files."""
    return int(os.environ.get("RANK", "0")) == 0


def _emit(text: str) -> None:
    if _is_root():
        print(text)


# ---------------------------------------------------------------------------
# file checks
# ---------------------------------------------------------------------------
def validate_file_readable(filepath: str, description: str) -> None:
    if not os.path.isfile(filepath):
        sys.exit(f"Error: {description} file '{filepath}' does not exist or is not a file.")
    if not os.access(filepath, os.R_OK):
        sys.exit(f"Error: {description} file '{filepath}' is not readable.")


def validate_file_writable(filepath: str, description: str) -> None:
    folder = os.path.dirname(filepath) or "."
    if os.path.exists(filepath):
        if not os.access(filepath, os.W_OK):
            sys.exit(f"Error: {description} file '{filepath}' is not writable.")
    elif not os.access(folder, os.W_OK):
        sys.exit(f"Error: Directory '{folder}' is not writable to create {description} file '{filepath}'.")


# ---------------------------------------------------------------------------
# arguments
# ---------------------------------------------------------------------------
def parse_arguments(args: Optional[List[str]] = None) -> argparse.Namespace:
    parser = argparse.ArgumentParser(prog="Biosequence project")
    parser.add_argument("-t", "--task", required=True, help="Task to execute")
    parser.add_argument("-g", "--genomefile", help="Genome FASTA file (multiple records)")
    parser.add_argument("-k", "--kmer-size", type=int, help="Length of k-mers")
    parser.add_argument("-r", "--referencefile", help="KDB file (input/output)")
    parser.add_argument("-a", "--alignfile", help="aln file. Can be either input or name for output file")
    parser.add_argument("--reads", help="FASTQ reads file")
    parser.add_argument("-m", "--unique-threshold", type=int, help="unique k-mer threshold")
    parser.add_argument("-p", "--ambiguous-threhold", type=int, help="ambiguous k-mer threshold")
    parser.add_argument("--reverse-complement", action="store_true")
    parser.add_argument("--min-read-quality", type=int, default=None)
    parser.add_argument("--min-kmer-quality", type=int, default=None)
    parser.add_argument("--max-genomes", type=int, default=None)
    parser.add_argument("--filter-similar", action="store_true")
    parser.add_argument("--similarity-threshold", type=float)
    return parser.parse_args(args)


# ---------------------------------------------------------------------------
# building blocks (names kept from the reference's module surface)
# ---------------------------------------------------------------------------
def create_reference(fasta_file: str, kmer_size: int, filter_similar: bool = False,
                     similarity_threshold: float = 0.95) -> KmerReference:
    return KmerReference(kmer_size, FASTAFile(fasta_file).container, filter_similar=filter_similar,
                         similarity_threshold=similarity_threshold)


def create_reference_and_save_it(fasta_file: str, kmer_size: int, reference_file: str, filter_similar: bool = False,
                                 similarity_threshold: float = 0.95) -> None:
    reference = create_reference(fasta_file, kmer_size, filter_similar, similarity_threshold)
    if _is_root():
        reference.save(reference_file)


def _load(kind, path: str):
    try:
        return kind.load(path)
    except gzip.BadGzipFile:
        sys.exit(BAD_FORMAT)


def dump_reference(kmer_reference: KmerReference) -> None:
    # the text of json.dumps(kmer_reference.get_summary(), indent=4) (main.py:127), written without the nested dicts
    _emit(kmer_reference.summary_json(indent=4))


def dump_reference_file(reference_file: str) -> None:
    dump_reference(_load(KmerReference, reference_file))


def build_reference_and_dump_from_file(fasta_file: str, kmer_size: int, filter_similar: bool = False,
                                       similarity_threshold: float = 0.95) -> None:
    dump_reference(create_reference(fasta_file, kmer_size, filter_similar, similarity_threshold))


def create_alignment_from_reference(kmer_reference: KmerReference, reads_file: str, m: int, p: int, min_read_quality,
                                    min_kmer_quality, max_genomes) -> PseudoAlignment:
    alignment = PseudoAlignment(kmer_reference)
    alignment.align_reads_from_container(FASTAQFile(reads_file).container, m, p, min_read_quality, min_kmer_quality,
                                         max_genomes)
    return alignment


def create_alignment_file_from_reference(kmer_reference: KmerReference, reads_file: str, align_file: str, m: int, p: int,
                                         min_read_quality, min_kmer_quality, max_genomes) -> None:
    alignment = create_alignment_from_reference(kmer_reference, reads_file, m, p, min_read_quality, min_kmer_quality, max_genomes)
    if _is_root():
        alignment.save(align_file)


def create_alignment_from_reference_file(reference_file: str, reads_file: str, align_file: str, m: int, p: int,
                                         min_read_quality, min_kmer_quality, max_genomes) -> None:
    create_alignment_file_from_reference(_load(KmerReference, reference_file), reads_file, align_file, m, p,
                                         min_read_quality, min_kmer_quality, max_genomes)


def build_reference_and_create_alignment_file(fasta_file: str, kmer_size: int, reads_file: str, align_file: str, m: int,
                                              p: int, min_read_quality, min_kmer_quality, max_genomes,
                                              filter_similar: bool = False, similarity_threshold: float = 0.95) -> None:
    create_alignment_file_from_reference(create_reference(fasta_file, kmer_size, filter_similar, similarity_threshold),
                                         reads_file, align_file, m, p, min_read_quality, min_kmer_quality, max_genomes)


def dump_alignment_file(align_file: str) -> None:
    _emit(json.dumps(_load(PseudoAlignment, align_file).get_summary(), indent=4))


def dump_alignment_from_reference(reference_file: str, reads_file: str, m: int, p: int, min_read_quality,
                                  min_kmer_quality, max_genomes) -> None:
    alignment = create_alignment_from_reference(_load(KmerReference, reference_file), reads_file, m, p,
                                                min_read_quality, min_kmer_quality, max_genomes)
    _emit(json.dumps(alignment.get_summary(), indent=4))


def build_reference_align_and_dump(fasta_file: str, kmer_size: int, reads_file: str, m: int, p: int, min_read_quality,
                                   min_kmer_quality, max_genomes, filter_similar: bool = False,
                                   similarity_threshold: float = 0.95) -> None:
    alignment = create_alignment_from_reference(create_reference(fasta_file, kmer_size, filter_similar, similarity_threshold),
                                                reads_file, m, p, min_read_quality, min_kmer_quality, max_genomes)
    _emit(json.dumps(alignment.get_summary(), indent=4))


# ---------------------------------------------------------------------------
# entry point
# ---------------------------------------------------------------------------
def _check_flag_combination(a: argparse.Namespace) -> None:
    read_side = (a.reads or a.alignfile or a.unique_threshold or a.ambiguous_threhold or a.min_read_quality
                 or a.min_kmer_quality or a.max_genomes)
    from_fasta = a.genomefile and a.kmer_size and a.reads
    if a.task == "reference":
        if read_side:
            sys.exit("Error: For task 'reference', only -g, -k, -r, --filter-similar, and --similarity-threshold are allowed.")
    elif a.task == "dumpref":
        if read_side:
            sys.exit("Error: For task 'dumpref', only -r or (-g and -k) with --filter-similar and --similarity-threshold are allowed.")
    elif a.task == "align":
        if not ((a.referencefile and a.reads and a.alignfile) or (from_fasta and a.alignfile)):
            sys.exit("Error: For task 'align', provide either -r (reference file) or -g and -k (genome file and kmer size) along with --reads and -a.")
    elif a.task == "dumpalign":
        if not ((a.referencefile and a.reads) or from_fasta or a.alignfile):
            sys.exit("Error: For task 'dumpalign', provide either -r and --reads, or -g, -k, and --reads, or -a.")
    else:
        sys.exit("Error: Unsupported task.")


def _run(a: argparse.Namespace) -> None:
    m, p = a.unique_threshold, a.ambiguous_threhold
    filters = (a.min_read_quality, a.min_kmer_quality, a.max_genomes)
    if a.task == "reference":
        validate_file_readable(a.genomefile, "Genome FASTA")
        validate_file_writable(a.referencefile, "Reference database output")
        create_reference_and_save_it(a.genomefile, a.kmer_size, a.referencefile, a.filter_similar, a.similarity_threshold)
    elif a.task == "dumpref":
        if a.referencefile:
            validate_file_readable(a.referencefile, "Reference database")
            dump_reference_file(a.referencefile)
        elif a.genomefile and a.kmer_size:
            validate_file_readable(a.genomefile, "Genome FASTA")
            build_reference_and_dump_from_file(a.genomefile, a.kmer_size, a.filter_similar, a.similarity_threshold)
    elif a.task == "align":
        validate_file_readable(a.reads, "FASTQ reads")
        validate_file_writable(a.alignfile, "Alignment output")
        if a.referencefile and a.reads and a.alignfile:
            validate_file_readable(a.referencefile, "Reference database")
            create_alignment_from_reference_file(a.referencefile, a.reads, a.alignfile, m, p, *filters)
        else:
            validate_file_readable(a.genomefile, "Genome FASTA")
            build_reference_and_create_alignment_file(a.genomefile, a.kmer_size, a.reads, a.alignfile, m, p, *filters,
                                                      a.filter_similar, a.similarity_threshold)
    elif a.task == "dumpalign":
        if a.referencefile and a.reads:
            validate_file_readable(a.reads, "FASTQ reads")
            dump_alignment_from_reference(a.referencefile, a.reads, m, p, *filters)
        elif a.genomefile and a.kmer_size and a.reads:
            validate_file_readable(a.reads, "FASTQ reads")
            validate_file_readable(a.genomefile, "Genome FASTA")
            build_reference_align_and_dump(a.genomefile, a.kmer_size, a.reads, m, p, *filters, a.filter_similar,
                                           a.similarity_threshold)
        elif a.alignfile:
            validate_file_readable(a.alignfile, "Alignment output")
            dump_alignment_file(a.alignfile)
        else:
            sys.exit("Error: Provide either -g and -k with --reads, or -r with --reads, or -a.")


def main() -> None:
    args = parse_arguments()
    _check_flag_combination(args)
    # defaults are filled in after validation with truthiness tests, so "-m 0" also becomes 1 (main.py:337-342)
    args.unique_threshold = args.unique_threshold or DEFAULT_UNIQUE_THRESHOLD
    args.ambiguous_threhold = args.ambiguous_threhold or DEFAULT_AMBIGUOUS_THRESHOLD
    args.similarity_threshold = args.similarity_threshold or DEFAULT_SIMILARITY_THRESHOLD
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:   # one rank of a torchrun launch: attach to the others
        import multi_gpu
        multi_gpu.init()
    try:
        _run(args)
    except gzip.BadGzipFile:
        sys.exit(BAD_FORMAT)
    except (InvalidExtensionError, NoRecordsInDataFile, NotValidatingUniqueMapping, AddingExistingRead, ValueError) as err:
        sys.exit(err)


if __name__ == "__main__":
    main()
