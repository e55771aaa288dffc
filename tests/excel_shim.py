"""In-memory stand-in for the Excel layer (openpyxl is not installed here): pd.read_excel / pd.ExcelFile / pd.ExcelWriter /
DataFrame.to_excel backed by a dict of frames, installed around a call.  TEST INFRASTRUCTURE -- used by the golden
generators (around the UNMODIFIED reference) and by the tests (around the drop-in), so both see the same "workbooks"."""
from __future__ import annotations

import contextlib
from pathlib import Path

import pandas as pd

BOOK: dict[str, dict[str, pd.DataFrame]] = {}


class _Writer:
    def __init__(self, path, *a, **k):
        self.path = str(path); BOOK[self.path] = {}
        Path(self.path).parent.mkdir(parents=True, exist_ok=True)
        Path(self.path).touch()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


class _File:
    def __init__(self, path, *a, **k):
        self.path = str(path)
        self.sheet_names = list(BOOK[self.path])


def _read_excel(path, sheet_name=None, **k):
    book = BOOK[str(path.path if isinstance(path, _File) else path)]
    return book[sheet_name or next(iter(book))].copy()


def _to_excel(self, target, sheet_name="Sheet1", index=True, **k):
    if isinstance(target, _Writer):
        BOOK[target.path][sheet_name] = self.copy()
    else:
        BOOK.setdefault(str(target), {})[sheet_name] = self.copy()
        Path(str(target)).parent.mkdir(parents=True, exist_ok=True)
        Path(str(target)).touch()


def put_book(path, sheets: dict):
    """Register a workbook and create an empty file at its path (the callers check os.path.exists)."""
    BOOK[str(path)] = {k: v.copy() for k, v in sheets.items()}
    Path(str(path)).parent.mkdir(parents=True, exist_ok=True)
    Path(str(path)).touch()


@contextlib.contextmanager
def installed():
    saved = (pd.read_excel, pd.ExcelWriter, pd.ExcelFile, pd.DataFrame.to_excel)
    pd.read_excel, pd.ExcelWriter, pd.ExcelFile, pd.DataFrame.to_excel = _read_excel, _Writer, _File, _to_excel
    try:
        yield BOOK
    finally:
        pd.read_excel, pd.ExcelWriter, pd.ExcelFile, pd.DataFrame.to_excel = saved
