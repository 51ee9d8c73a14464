"""
records.py -- host-side FASTA / FASTQ record model (argument types of the hot-path API).

API-compatible with the reference's records module (/root/reference/src/records.py):
Section / SectionSpecification / Record / RecordContainer / FASTARecordContainer /
FASTAQRecordContainer and the four exceptions, with the same acceptance rules
(records.py:141-199, 212-302).  Text parsing is outside the accelerated path; what
changes here is only bookkeeping that does not scale: the reference tracks parsed
input with one Python int per character (records.py:170-181), this version tracks
the matched spans as intervals.

Ingest fast path: canonical ASCII text is parsed natively (csrc/ingest.cpp) straight into the
packed arrays the device path consumes; Record objects are then created lazily, only if somebody iterates the
container.  Anything the native parser does not recognise as canonical goes through the regular expression below, so
unusual inputs keep the reference's acceptance rules, exception types, messages and precedence.
"""
import re
from collections import namedtuple
from typing import Iterator, List, Optional, Sequence, Tuple

import constants

UNTIL_NEXT_HEADER_OR_EOF = r"(?=(?=\r?\n{section_header})|(?=(?:\r?\n)?\Z))"
UNPARSED_SNIPPET_LEN = 20
NATIVE_INGEST = True   # tests switch this off to compare the native parser with the regular expression


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
        self._packed = None          # natively parsed batch (see _native.parse_records_native), or None
        self._packed_pending = False  # True while the Records of the packed batch have not been materialised

    NATIVE_KIND: Optional[bool] = None   # True = FASTQ, False = FASTA, None = no native parser for this container

    def _try_native(self, data: str) -> bool:
        if not NATIVE_INGEST or type(self).NATIVE_KIND is None or self._records or self._packed is not None or not isinstance(data, str):
            return False
        if not data.isascii():
            return False
        try:
            import _native as nat
            packed = nat.parse_records_native(data, bool(type(self).NATIVE_KIND))
        except ImportError:     # the library has not been built: the regex path below needs nothing native
            return False
        if packed is None:
            return False
        self._packed, self._packed_pending = packed, True
        return True

    def _materialize(self) -> None:
        """Creates the Record objects of a natively parsed batch (lazily: the hot path never needs them)."""
        if not self._packed_pending:
            return
        self._packed_pending = False
        import _native as nat
        pk = self._packed
        raw, off = pk["raw"], pk["off"].tolist()
        seq = pk["seq"].tobytes().decode("ascii")
        qual = pk["qual"].tobytes().decode("ascii") if pk["qual"] is not None else None
        specs = type(self).SECTION_SPECIFICATIONS
        for i, (b, l) in enumerate(zip(pk["name_beg"].tolist(), pk["name_len"].tolist())):
            name = nat.parsed_text(raw, b, l)
            if qual is None:
                fields = (name, seq[off[i]:off[i + 1]])
            else:
                pb, pl = int(pk["plus_beg"][i]), int(pk["plus_len"][i])
                fields = (name, seq[off[i]:off[i + 1]], nat.parsed_text(raw, pb, pl).strip(), qual[off[i]:off[i + 1]])
            self._records.append(Record([Section(spec.section_name, f) for spec, f in zip(specs, fields)]))
            for spec, f in zip(specs, fields):
                if spec.is_unique_index:
                    self._unique_index_values.add(f)

    def packed_batch(self):
        """The natively parsed arrays when this container holds exactly that batch, else None."""
        if self._packed is not None and (self._packed_pending or len(self._records) == self._packed["n"]):
            return self._packed
        return None

    def __getstate__(self):
        self._materialize()
        state = dict(self.__dict__)
        state["_packed"], state["_packed_pending"] = None, False
        return state

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
        if self._try_native(data):
            return
        self._materialize()
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
        self._materialize()
        return iter(self._records)

    def __len__(self) -> int:
        return self._packed["n"] if self._packed_pending else len(self._records)


class FASTARecordContainer(RecordContainer):
    """'>' description line, then the genome over [ACGTN] with whitespace stripped."""

    NATIVE_KIND = False

    SECTION_SPECIFICATIONS = (
        SectionSpecification("description", ">", True, r"\S\t ", "", False),
        SectionSpecification("genome", "", True, constants.NUCLEOTIDES_CHARS, r"\s", False),
    )


class FASTAQRecordContainer(RecordContainer):
    """'@' identifier (unique), sequence over [ACGT], '+' line, quality over ASCII 33..126 of equal length."""

    NATIVE_KIND = True

    SECTION_SPECIFICATIONS = (
        SectionSpecification("identifier", "@", True, r"\S\t ", "", True),
        SectionSpecification("sequence", "", True, re.escape(constants.REAL_NUCLEOTIDES_CHARS), "", False),
        SectionSpecification("space", "+", False, ".", "", False),
        SectionSpecification("quality_sequence", "", True, re.escape("".join(constants.PHRED33_SCORES)), "", False),
    )

    def parse_records(self, data: str) -> None:
        super().parse_records(data)
        if self._packed_pending:
            return      # the native parser only accepts records whose two lengths agree
        for number, record in enumerate(self, start=1):
            n_seq, n_qual = len(record["sequence"]), len(record["quality_sequence"])
            if n_seq != n_qual:
                raise InvalidRecordData(f"Mismatch in record {number} between nucleotide length: {n_seq} "
                                        f"and PHRED section lengths: {n_qual}")
