"""Dict-backed stand-in for ``pandas.HDFStore`` (PyTables is not installed in this image).  TEST INFRASTRUCTURE ONLY.

It implements the slice of the HDFStore API that GPSat's results path uses -- ``append`` (table format: rows are
appended, the schema is fixed by the first append), ``get``, ``select`` with the ``"col == value"`` where-strings of
``LocalExpertOI._read_params_from_file`` (GPSat/local_experts.py:652-660), ``keys``, ``in``, ``get_storer(k).attrs``,
context-manager use -- over a process-global ``{path: {table: DataFrame}}``, creates an (empty) file at ``path`` so
``os.path.exists`` behaves, and RECORDS the keyword arguments of every ``append`` (``min_itemsize``, ``data_columns``,
``index``) so tests can assert schema parity between the reference's writer (GPSat/local_experts.py:499-550,
GPSat/utils.py:1195-1254) and gpsat_b200's.

``install()`` patches ``pd.HDFStore`` / ``pd.io.pytables.HDFStore`` (isinstance checks in GPSat/dataloader.py:1161 keep
working) and ``pd.read_hdf``; ``uninstall()`` restores them.
"""
import os
import re

import numpy as np
import pandas as pd

FILES = {}          # abspath -> {"tables": {name: DataFrame}, "attrs": {name: dict}, "appends": [(name, kwargs, nrows)]}
_REAL = {}


def _entry(path, create):
    p = os.path.abspath(path)
    if p not in FILES:
        if not create:
            raise OSError(f"File {path} does not exist")
        FILES[p] = {"tables": {}, "attrs": {}, "appends": []}
        os.makedirs(os.path.dirname(p), exist_ok=True)
        open(p, "ab").close()
    return FILES[p]


class _Storer:
    def __init__(self, attrs):
        self.attrs = attrs


class FakeHDFStore:
    def __init__(self, path, mode="a", **kwargs):
        self._path = path
        self._mode = mode
        self._e = _entry(path, create=(mode != "r"))
        self.is_open = True

    # context manager
    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
        return False

    def close(self):
        self.is_open = False

    @staticmethod
    def _k(key):
        return key.lstrip("/")

    def keys(self):
        return ["/" + k for k in self._e["tables"]]

    def __contains__(self, key):
        return self._k(key) in self._e["tables"]

    def __iter__(self):
        return iter(self.keys())

    def get(self, key):
        k = self._k(key)
        if k not in self._e["tables"]:
            raise KeyError(f"No object named {key} in the file")
        return self._e["tables"][k].copy()

    __getitem__ = get

    def get_storer(self, key):
        k = self._k(key)
        if k not in self._e["tables"]:
            raise KeyError(f"No object named {key} in the file")
        return _Storer(self._e["attrs"].setdefault(k, {}))

    def append(self, key, value, **kwargs):
        assert self._mode != "r", "store opened read-only"
        k = self._k(key)
        self._e["appends"].append((k, dict(kwargs), len(value)))
        cur = self._e["tables"].get(k)
        if cur is None:
            self._e["tables"][k] = value.copy()
            return
        # table format: the schema is fixed by the first append
        if list(cur.columns) != list(value.columns) or list(cur.index.names) != list(value.index.names):
            raise ValueError(f"cannot match existing table structure for [{','.join(map(str, value.columns))}] "
                             f"on appending data to {k}")
        mi = kwargs.get("min_itemsize") or {}
        for c in value.columns:
            if value[c].dtype == object and len(value):
                width = int(value[c].astype(str).str.len().max())
                first = self._e["attrs"].setdefault(k, {}).setdefault("_itemsize", {})
                limit = first.get(c)
                if limit is not None and width > limit:
                    raise ValueError(f"Trying to store a string with len [{width}] in [{c}] column but this column "
                                     f"has a limit of [{limit}]!")
        self._e["tables"][k] = pd.concat([cur, value], axis=0)

    def put(self, key, value, **kwargs):
        self._e["tables"][self._k(key)] = value.copy()

    def remove(self, key):
        self._e["tables"].pop(self._k(key))

    def select(self, key, where=None, columns=None, **kwargs):
        df = self.get(key)
        if where is not None:
            if isinstance(where, str):
                where = [where]
            flat = df.reset_index()
            m = np.ones(len(flat), dtype=bool)
            for w in where:
                mt = re.match(r"^\s*([A-Za-z_]\w*)\s*(==|>=|<=|!=|>|<|=)\s*(.+?)\s*$", w)
                assert mt, f"where string not understood by the fake store: {w}"
                col, comp, val = mt.groups()
                comp = "==" if comp == "=" else comp
                v = val.strip("'\"")
                colv = flat[col].values
                if np.issubdtype(colv.dtype, np.datetime64):
                    v = np.datetime64(v)
                elif np.issubdtype(colv.dtype, np.number):
                    v = float(v)
                m &= {"==": np.equal, ">=": np.greater_equal, "<=": np.less_equal, ">": np.greater, "<": np.less,
                      "!=": np.not_equal}[comp](colv, v)
            df = df.loc[m]
        if columns is not None:
            df = df[columns]
        return df


def read_hdf(path_or_buf, key=None, **kwargs):
    with FakeHDFStore(path_or_buf, mode="r") as st:
        return st.select(key, **{k: v for k, v in kwargs.items() if k in ("where", "columns")})


def install():
    if not _REAL:
        _REAL["HDFStore"] = pd.HDFStore
        _REAL["io"] = pd.io.pytables.HDFStore
        _REAL["read_hdf"] = pd.read_hdf
    pd.HDFStore = FakeHDFStore
    pd.io.pytables.HDFStore = FakeHDFStore
    pd.read_hdf = read_hdf


def uninstall():
    if _REAL:
        pd.HDFStore = _REAL["HDFStore"]
        pd.io.pytables.HDFStore = _REAL["io"]
        pd.read_hdf = _REAL["read_hdf"]
    FILES.clear()


def tables(path):
    return FILES[os.path.abspath(path)]["tables"]


def appends(path):
    return FILES[os.path.abspath(path)]["appends"]
