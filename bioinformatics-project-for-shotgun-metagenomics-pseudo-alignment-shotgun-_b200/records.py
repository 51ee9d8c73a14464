"""
records.py -- host-side FASTA / FASTQ record model (argument types of the hot-path API).

API-compatible with the reference's records module (/root/reference/src/records.py):
Section / SectionSpecification / Record / RecordContainer / FASTARecordContainer /
FASTAQRecordContainer and the four exceptions, with the same acceptance rules
(records.py:141-199, 212-302).  Text parsing is outside the accelerated path; what
changes here is only bookkeeping that does not scale: the reference tracks parsed
input with one Python int per character (records.py:170-181), this version tracks
the matched spans as intervals.
"""
import re
from collections import namedtuple
from typing import Iterator, List, Optional, Sequence, Tuple

import constants

UNTIL_NEXT_HEADER_OR_EOF = r"(?=(?=\r?\n{section_header})|(?=(?:\r?\n)?\Z))"
UNPARSED_SNIPPET_LEN = 20


class NoRecordsInData(Exception):
    def __init__(self, message: str = "No valid records found in the data.") -> None:
        super().__init__(message)


class InvalidRecordData(Exception):
    def __init__(self, message: str = "") -> None:
        super().__init__(message)


class DuplicateRecordError(Exception):
    def __init__(self, message: str = "Duplicate records found for the unique index.") -> None:
        super().__init__(message)


class UnparsedDataError(Exception):
    def __init__(self, message: str = "Unparsed data found in the input.") -> None:
        super().__init__(message)


Section = namedtuple("Section", ["name", "data"])
SectionSpecification = namedtuple(
    "SectionSpecification",
    ["section_name", "section_header", "must_have_data", "section_legal_chars", "chars_to_remove", "is_unique_index"],
)


class Record(object):
    """One parsed record: ordered named sections; `identifier` is the first section's data.

    Records are used as dict keys by identity, exactly like the reference's (no __eq__/__hash__)."""

    def __init__(self, sections: Sequence[Section]) -> None:
        if len(sections) == 0:
            raise InvalidRecordData("The data given to construct record has no sections.")
        self.identifier: str = sections[0].data
        table = {}
        for name, data in sections:
            if name in table:
                raise InvalidRecordData(f"Section header: {name} has appeared twice in the given data.")
            table[name] = data
        self.__sections = table

    def __getitem__(self, key: str) -> str:
        return self.__sections[key]

    def __str__(self) -> str:
        return "\n".join(f"{name}: {data}" for name, data in self.__sections.items())

    __repr__ = __str__


def _first_unparsed(data: str, spans: List[Tuple[int, int]]) -> Optional[int]:
    """Index of the first non-whitespace character outside every matched span (spans are ascending)."""
    cursor = 0
    for start, end in spans + [(len(data), len(data))]:
        gap = data[cursor:start]
        stripped = gap.lstrip()
        if stripped:
            return cursor + (len(gap) - len(stripped))
        cursor = max(cursor, end)
    return None


class RecordContainer(object):
    """Base container: subclasses describe a record as a tuple of SectionSpecification."""

    SECTION_SPECIFICATIONS: Tuple[SectionSpecification, ...]

    def __init__(self) -> None:
        if getattr(type(self), "SECTION_SPECIFICATIONS", None) is None:
            raise NotImplementedError("SECTION_SPECIFICATIONS must be defined.")
        self.__re_pattern: str = ""
        self.create_record_re_string()
        self._unique_index_values = set()
        self._records: List[Record] = []

    def create_record_re_string(self) -> None:
        """One lazy capture group per section, terminated by a look-ahead for the next record or the end."""
        specs = type(self).SECTION_SPECIFICATIONS
        pieces = []
        for position, spec in enumerate(specs):
            lead = "^" if position == 0 else r"\r?\n"
            repeat = "+?" if spec.must_have_data else "*?"
            pieces.append(f"{lead}{re.escape(spec.section_header)}((?:[{spec.section_legal_chars}{spec.chars_to_remove}]){repeat})")
        pieces.append(UNTIL_NEXT_HEADER_OR_EOF.format(section_header=re.escape(specs[0].section_header)))
        self.__re_pattern = "".join(pieces)

    def parse_records(self, data: str) -> None:
        spans: List[Tuple[int, int]] = []
        for found in re.finditer(self.__re_pattern, data, flags=re.MULTILINE):
            groups = found.groups()
            if any(groups):
                spans.append(found.span())
                self.create_record(groups)
        if not self._records:
            raise NoRecordsInData
        stray = _first_unparsed(data, spans)
        if stray is not None:
            raise UnparsedDataError(f"Unparsed data found at index {stray}: {data[stray:stray + UNPARSED_SNIPPET_LEN]}...")

    def create_record(self, record_match_groups) -> None:
        sections = []
        for spec, raw in zip(type(self).SECTION_SPECIFICATIONS, record_match_groups):
            cleaned = re.sub(spec.chars_to_remove, "", raw or "").strip()
            sections.append(Section(spec.section_name, cleaned))
            if spec.is_unique_index:
                if cleaned in self._unique_index_values:
                    raise DuplicateRecordError(f"Duplicate record found with unique index: {cleaned}")
                self._unique_index_values.add(cleaned)
        self._records.append(Record(sections))

    def __iter__(self) -> Iterator[Record]:
        return iter(self._records)

    def __len__(self) -> int:
        return len(self._records)


class FASTARecordContainer(RecordContainer):
    """'>' description line, then the genome over [ACGTN] with whitespace stripped."""

    SECTION_SPECIFICATIONS = (
        SectionSpecification("description", ">", True, r"\S\t ", "", False),
        SectionSpecification("genome", "", True, constants.NUCLEOTIDES_CHARS, r"\s", False),
    )


class FASTAQRecordContainer(RecordContainer):
    """'@' identifier (unique), sequence over [ACGT], '+' line, quality over ASCII 33..126 of equal length."""

    SECTION_SPECIFICATIONS = (
        SectionSpecification("identifier", "@", True, r"\S\t ", "", True),
        SectionSpecification("sequence", "", True, re.escape(constants.REAL_NUCLEOTIDES_CHARS), "", False),
        SectionSpecification("space", "+", False, ".", "", False),
        SectionSpecification("quality_sequence", "", True, re.escape("".join(constants.PHRED33_SCORES)), "", False),
    )

    def parse_records(self, data: str) -> None:
        super().parse_records(data)
        for number, record in enumerate(self, start=1):
            n_seq, n_qual = len(record["sequence"]), len(record["quality_sequence"])
            if n_seq != n_qual:
                raise InvalidRecordData(f"Mismatch in record {number} between nucleotide length: {n_seq} "
                                        f"and PHRED section lengths: {n_qual}")
