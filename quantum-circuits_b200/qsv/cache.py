"""On-disk cache of host-side artefacts that are expensive to rebuild in a fresh process: joint placements of
sharded circuits (seconds of search) and the cubins of pass-specialised kernels (NVRTC, ~0.3 s each).

Nothing here is on the device path; entries are keyed by a content digest of everything that determines them.
QSV_CACHE_DIR chooses the directory (default ~/.cache/qsv), QSV_DISK_CACHE=0 switches the cache off.  Writes are
atomic (temp file + rename), so concurrent ranks of one job may share the directory."""
import hashlib
import os
import pickle
import tempfile

VERSION = b"qsv-cache-2"


def enabled():
    return os.environ.get("QSV_DISK_CACHE", "1") != "0"


def directory():
    d = os.environ.get("QSV_CACHE_DIR") or os.path.join(os.path.expanduser("~"), ".cache", "qsv")
    return d


def digest(data):
    return hashlib.blake2b(VERSION + data, digest_size=16).hexdigest()


def _path(name):
    return os.path.join(directory(), name)


def load_bytes(name):
    if not enabled():
        return None
    try:
        with open(_path(name), "rb") as f:
            return f.read()
    except OSError:
        return None


def store_bytes(name, data):
    if not enabled():
        return
    try:
        d = directory()
        os.makedirs(d, exist_ok=True)
        fd, tmp = tempfile.mkstemp(dir=d, prefix=".tmp-")
        with os.fdopen(fd, "wb") as f:
            f.write(data)
        os.replace(tmp, _path(name))
    except OSError:
        pass                 # a read-only home directory only costs the cache


def load_pickle(name):
    raw = load_bytes(name)
    if raw is None:
        return None
    try:
        return pickle.loads(raw)
    except Exception:        # noqa: BLE001 - a truncated or foreign file is a miss
        return None


def store_pickle(name, value):
    try:
        store_bytes(name, pickle.dumps(value, protocol=4))
    except Exception:        # noqa: BLE001
        pass
