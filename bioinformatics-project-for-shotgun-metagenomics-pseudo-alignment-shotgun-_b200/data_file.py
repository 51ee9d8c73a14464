"""
data_file.py -- extension check + whole-file read (optionally gzip) into a record container.

Host-side shim with the reference's names and error behaviour
(/root/reference/src/data_file.py:39-158); file I/O is outside the accelerated path.
"""
import gzip
import pickle
from typing import Optional, Set

from records import FASTARecordContainer, FASTAQRecordContainer, NoRecordsInData, RecordContainer


class InvalidExtensionError(Exception):
    def __init__(self, message: str = "") -> None:
        super().__init__(message)


class NoRecordsInDataFile(Exception):
    def __init__(self, message: str = "") -> None:
        super().__init__(message)


class DataFile:
    EXTENSIONS: Optional[Set[str]] = None

    def __init__(self, file_path: str) -> None:
        extensions = type(self).EXTENSIONS
        if not extensions:
            raise NotImplementedError("EXTENSIONS must be defined.")
        if not file_path.endswith(tuple(extensions)):
            raise InvalidExtensionError(f"Invalid file extension. Expected one of {extensions}, got {file_path}")
        self.container: RecordContainer = self.get_container_type()
        self.parse_file(file_path)

    def get_container_type(self) -> RecordContainer:
        raise NotImplementedError("This method must be implemented in subclasses.")

    def parse_file(self, file_path: str) -> None:
        try:
            self.container.parse_records(self.load_file(file_path))
        except NoRecordsInData:
            raise NoRecordsInDataFile(f"No valid records found in file: {file_path}")

    def load_file(self, file_path: str) -> str:
        opener = gzip.open if file_path.endswith(".gz") else open
        with opener(file_path, "rt", encoding="utf-8") as handle:
            return handle.read()

    def dump(self, output_file: str) -> None:
        with open(output_file, "wb") as handle:
            pickle.dump(self.container, handle)


class FASTAFile(DataFile):
    EXTENSIONS = {".fa", ".fa.gz"}

    def get_container_type(self) -> FASTARecordContainer:
        return FASTARecordContainer()


class FASTAQFile(DataFile):
    EXTENSIONS = {".fq", ".fq.gz"}

    def get_container_type(self) -> FASTAQRecordContainer:
        return FASTAQRecordContainer()
