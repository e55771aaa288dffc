"""Frames this process has just read or written, kept so that the next pipeline step does not parse the file again.

The reference's steps talk to each other through CSV files (processor.py:158 -> :181, :213 -> :235, :310 -> :379):
every step re-reads what the previous one wrote.  The drop-in keeps the file contract -- every output file is
written, synchronously, byte for byte -- and additionally remembers, per path, the frame ``pd.read_csv`` WOULD
return for the file it just wrote or read, keyed by (absolute path, size, mtime_ns, inode, a hash of the file's first and last 4 KB).  A later ``_read_csv``
of the same, unchanged file gets that frame back (a shallow copy: pandas >= 3 is copy-on-write, so callers cannot
alter the cached columns).  A file changed by anyone else has another size / mtime and is read from disk.

Soundness: a written frame is only remembered when reading it back could not change it -- text columns must pass
``dyd_csv_roundtrip_check`` (no cell that would come back as NaN, every dtype-inference chunk stays text), numeric
and boolean numpy columns round-trip through ``repr`` by construction, the column names must be unique non-empty
strings.  tests/test_tablecache.py holds every cached frame against ``pd.read_csv`` of the file.

``extras`` carries step-specific by-products of the frame (step 4 leaves the boxes of the new column for step 5).
Bounded by DYD_TABLE_CACHE_MB (default 4096) with LRU eviction; DYD_TABLE_CACHE=0 switches it off.
"""
from __future__ import annotations

import os
import threading
from collections import OrderedDict

_LOCK = threading.Lock()
_ENTRIES: "OrderedDict[tuple, Entry]" = OrderedDict()
_BYTES = 0
STATS = {"hits": 0, "misses": 0, "stores": 0, "declined": 0}


class Entry:
    __slots__ = ("frame", "extras", "nbytes", "clean", "lazy")

    def __init__(self, frame, extras, nbytes, clean, lazy=None):
        self.frame, self.extras, self.nbytes, self.clean, self.lazy = frame, extras or {}, nbytes, clean, lazy


def enabled() -> bool:
    if os.environ.get("DYD_TABLE_CACHE", "1") == "0":
        return False
    from . import native
    return native.enabled() and native._pandas_infers_arrow_str()


def _limit() -> int:
    return int(os.environ.get("DYD_TABLE_CACHE_MB", "4096")) << 20


def _key(path):
    """(absolute path, size, mtime_ns, inode, hash of the first and last 4 KB).  The content probe costs two small reads and
    closes the one gap stat() leaves: a same-size rewrite by somebody else inside one tick of the file system's clock."""
    try:
        with open(path, "rb") as f:
            st = os.fstat(f.fileno())
            head = f.read(4096)
            tail = b""
            if st.st_size > 8192:
                f.seek(st.st_size - 4096)
                tail = f.read(4096)
            elif st.st_size > 4096:
                tail = f.read()
    except OSError:
        return None
    return (os.path.abspath(os.fspath(path)), st.st_size, st.st_mtime_ns, st.st_ino, hash(head), hash(tail))


def _frame_bytes(df) -> int:
    total = 0
    for c in df.columns:
        arr = df[c].array
        pa_arr = getattr(arr, "_pa_array", None)
        total += int(pa_arr.nbytes) if pa_arr is not None else int(getattr(arr, "nbytes", 0))
    return total


def clear() -> None:
    global _BYTES
    with _LOCK:
        _ENTRIES.clear()
        _BYTES = 0


def forget(path) -> None:
    global _BYTES
    ap = os.path.abspath(os.fspath(path))
    with _LOCK:
        for k in [k for k in _ENTRIES if k[0] == ap]:
            _BYTES -= _ENTRIES.pop(k).nbytes


def get(path):
    """Entry of an unchanged file this process read or wrote, else None."""
    if not enabled():
        return None
    k = _key(path)
    with _LOCK:
        e = _ENTRIES.get(k) if k else None
        if e is None:
            STATS["misses"] += 1
            return None
        _ENTRIES.move_to_end(k)
        STATS["hits"] += 1
    if e.lazy is not None:                          # (parent frame, row numbers, check): the rows are gathered on first use
        parent, rows, check = e.lazy
        frame = parent.iloc[rows].reset_index(drop=True)
        if not check(frame):
            forget(path); declined()
            return None
        e.frame, e.lazy = frame, None
    return e


def put(path, frame, extras=None, clean=True, lazy=None) -> None:
    """Remember `frame` (RangeIndex, the dtypes pd.read_csv would infer) as the content of the file at `path` as it is on
    disk NOW.  `lazy` = (parent frame, row numbers, round-trip check) instead of a frame: gathered and checked when somebody asks."""
    global _BYTES
    if not enabled():
        return
    k = _key(path)
    if k is None:
        return
    # a lazy entry keeps its parent frame alive: charge it (conservatively, even where another entry shares the columns)
    nbytes = _frame_bytes(frame) if frame is not None else (_frame_bytes(lazy[0]) if lazy is not None else 0)
    with _LOCK:
        for old in [o for o in _ENTRIES if o[0] == k[0]]:
            _BYTES -= _ENTRIES.pop(old).nbytes
        if nbytes > _limit():
            STATS["declined"] += 1
            return
        _ENTRIES[k] = Entry(frame, extras, nbytes, clean, lazy)
        _BYTES += nbytes
        STATS["stores"] += 1
        while _BYTES > _limit() and len(_ENTRIES) > 1:
            _, ev = _ENTRIES.popitem(last=False)
            _BYTES -= ev.nbytes


def declined() -> None:
    STATS["declined"] += 1
