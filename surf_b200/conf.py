"""Tiny HOCON-subset reader for SuRF's ``confs/*.conf`` files.

The reference parses its configs with pyhocon (runner.py:35) which is not a
dependency here.  Only the subset the shipped confs use is supported: nested
``name { ... }`` blocks (with or without a space before the brace), ``k = v``
and ``k : v`` pairs, ``[a, b, ...]`` lists (possibly nested / multi-line),
``#`` and ``//`` comments, numbers, ``True/False/true/false``, quoted strings
and unquoted strings (e.g. ``<DTU path>`` or ``datasets/dtu_split/train.txt``).

``ConfigTree`` mirrors the accessors the reference calls on pyhocon trees
(``get_int/get_float/get_list/get_bool/get_string/get``, ``tree["a.b"]``,
``"k" in tree`` and ``**tree`` expansion, see implicit_surface.py:54-62).
"""
from __future__ import annotations

import re
_MISSING = object()


class ConfigMissingException(KeyError):
    pass


class ConfigTree(dict):
    """Ordered nested mapping with dotted-path access."""

    def _walk(self, key):
        cur = self
        for part in str(key).split("."):
            if not isinstance(cur, dict) or not dict.__contains__(cur, part):
                raise ConfigMissingException("No configuration setting found for key " + str(key))
            cur = dict.__getitem__(cur, part)
        return cur

    def __getitem__(self, key):
        return self._walk(key)

    def __contains__(self, key):
        try:
            self._walk(key)
            return True
        except ConfigMissingException:
            return False

    def get(self, key, default=_MISSING):
        try:
            return self._walk(key)
        except ConfigMissingException:
            if default is _MISSING:
                raise
            return default

    def get_int(self, key, default=_MISSING):
        v = self.get(key, default)
        return v if v is default and default is not _MISSING else int(v)

    def get_float(self, key, default=_MISSING):
        v = self.get(key, default)
        return v if v is default and default is not _MISSING else float(v)

    def get_bool(self, key, default=_MISSING):
        v = self.get(key, default)
        if v is default and default is not _MISSING:
            return v
        if isinstance(v, str):
            return v.strip().lower() in ("true", "yes", "on", "1")
        return bool(v)

    def get_string(self, key, default=_MISSING):
        v = self.get(key, default)
        return v if v is default and default is not _MISSING else str(v)

    def get_list(self, key, default=_MISSING):
        v = self.get(key, default)
        if v is default and default is not _MISSING:
            return v
        if not isinstance(v, list):
            raise TypeError("%s is not a list" % key)
        return list(v)

    def get_config(self, key, default=_MISSING):
        v = self.get(key, default)
        if v is not default and not isinstance(v, ConfigTree):
            raise TypeError("%s is not a config block" % key)
        return v

    def put(self, key, value):
        parts = str(key).split(".")
        cur = self
        for part in parts[:-1]:
            if not dict.__contains__(cur, part):
                dict.__setitem__(cur, part, ConfigTree())
            cur = dict.__getitem__(cur, part)
        dict.__setitem__(cur, parts[-1], value)

    @staticmethod
    def from_dict(d):
        t = ConfigTree()
        for k, v in d.items():
            dict.__setitem__(t, k, ConfigTree.from_dict(v) if isinstance(v, dict) else v)
        return t


_TOKEN = re.compile(
    r"""\s*(?:
        (?P<comment>(?:\#|//)[^\n]*) |
        (?P<brace>[{}\[\],=:]) |
        (?P<dq>"(?:[^"\\]|\\.)*") |
        (?P<sq>'(?:[^'\\]|\\.)*') |
        (?P<nl>\n) |
        (?P<bare>[^\s{}\[\],=:\#"']+(?:[ \t]+[^\s{}\[\],=:\#"']+)*)
    )""",
    re.VERBOSE,
)


def _tokens(text):
    pos, n = 0, len(text)
    while pos < n:
        m = _TOKEN.match(text, pos)
        if m is None:
            if text[pos:].strip() == "":
                return
            raise ValueError("conf syntax error near: %r" % text[pos:pos + 40])
        pos = m.end()
        kind = m.lastgroup
        if kind == "comment":
            continue
        yield kind, m.group(kind)


def _scalar(kind, tok):
    if kind in ("dq", "sq"):
        return tok[1:-1]
    low = tok.lower()
    if low in ("true", "false"):
        return low == "true"
    if low in ("null", "none"):
        return None
    try:
        return int(tok)
    except ValueError:
        pass
    try:
        return float(tok)
    except ValueError:
        return tok


class _Parser:
    def __init__(self, text):
        self.toks = list(_tokens(text))
        self.i = 0

    def peek(self):
        return self.toks[self.i] if self.i < len(self.toks) else (None, None)

    def next(self):
        t = self.peek()
        self.i += 1
        return t

    def skip_nl(self, also_commas=False):
        while True:
            k, v = self.peek()
            if k == "nl" or (also_commas and k == "brace" and v == ","):
                self.i += 1
            else:
                return

    def parse_block(self, top):
        tree = ConfigTree()
        while True:
            self.skip_nl(also_commas=True)
            k, v = self.peek()
            if k is None:
                if not top:
                    raise ValueError("unterminated '{' block")
                return tree
            if k == "brace" and v == "}":
                if top:
                    raise ValueError("unexpected '}'")
                self.i += 1
                return tree
            if k not in ("bare", "dq", "sq"):
                raise ValueError("expected a key, got %r" % v)
            self.i += 1
            key = v[1:-1] if k in ("dq", "sq") else v
            k2, v2 = self.peek()
            if k2 == "brace" and v2 == "{":
                self.i += 1
                val = self.parse_block(False)
            elif k2 == "brace" and v2 in "=:":
                self.i += 1
                val = self.parse_value()
            else:
                raise ValueError("expected '=', ':' or '{' after key %r" % key)
            if key in tree and isinstance(tree.get(key), ConfigTree) and isinstance(val, ConfigTree):
                _merge(dict.__getitem__(tree, key), val)
            else:
                tree.put(key, val)

    def parse_value(self):
        k, v = self.next()
        if k == "brace" and v == "{":
            return self.parse_block(False)
        if k == "brace" and v == "[":
            out = []
            while True:
                self.skip_nl(also_commas=True)
                k2, v2 = self.peek()
                if k2 is None:
                    raise ValueError("unterminated '[' list")
                if k2 == "brace" and v2 == "]":
                    self.i += 1
                    return out
                out.append(self.parse_value())
        if k in ("bare", "dq", "sq"):
            return _scalar(k, v)
        raise ValueError("unexpected token %r in value position" % v)


def _merge(dst, src):
    for k, v in src.items():
        if k in dst and isinstance(dict.__getitem__(dst, k), ConfigTree) and isinstance(v, ConfigTree):
            _merge(dict.__getitem__(dst, k), v)
        else:
            dict.__setitem__(dst, k, v)


def parse_string(text) -> ConfigTree:
    return _Parser(text).parse_block(True)


def parse_file(path) -> ConfigTree:
    with open(path, "r") as f:
        return parse_string(f.read())


# The ``model.implicit_surface`` block shared by all five shipped confs
# (confs/surf.conf:87-123); used when no conf file is supplied.
DEFAULT_IMPLICIT_SURFACE_CONF = """
sdf_network {
    d_out = 129
    d_in = 3
    d_hidden = 128
    n_layers = 6
    skip_in = [3]
    multires = 4
    bias = 0.5
    scale = 1.0
    geometric_init = True
    weight_norm = True
    feat_channels = 28
    feat_multires = 0
}
color_network {
    d_feature = 16
}
variance_network {
    init_val = 0.3
}
render {
    n_samples = [64, 32, 24, 16]
    sample_ranges = [1.0, 0.4, 0.1, 0.01]
    n_depth = 256
    perturb = 1.0
}
"""


def default_implicit_surface_conf() -> ConfigTree:
    return parse_string(DEFAULT_IMPLICIT_SURFACE_CONF)
