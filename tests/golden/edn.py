"""Minimal EDN reader used ONLY to generate the golden fixtures in this directory.

It reads the subset of EDN that appears in the reference's rendered tutorial
(`doc/tutorial.md` result maps) and in `resources/simulator-devices.edn`:
maps, vectors, lists, sets, keywords, symbols, strings, chars, numbers, ratios,
nil/true/false, `#object[...]`/`#function[...]` tagged junk (returned as Tagged),
comments and the `#_` discard form.  Keywords become `":kw"` strings so that the
fixtures can be written as JSON.
"""
from __future__ import annotations

from fractions import Fraction


class Tagged:
    def __init__(self, tag, value):
        self.tag, self.value = tag, value

    def __repr__(self):
        return f"#{self.tag} {self.value!r}"


class Sym(str):
    pass


class EDNError(ValueError):
    pass


_WS = " \t\r\n,"
_DELIM = _WS + "()[]{}\";"


class _Reader:
    def __init__(self, text: str, pos: int = 0):
        self.s, self.i, self.n = text, pos, len(text)

    def skip(self):
        s = self.s
        while self.i < self.n:
            c = s[self.i]
            if c in _WS:
                self.i += 1
            elif c == ";":
                while self.i < self.n and s[self.i] != "\n":
                    self.i += 1
            else:
                break

    def read(self):
        self.skip()
        if self.i >= self.n:
            raise EDNError("eof")
        c = self.s[self.i]
        if c == "{":
            self.i += 1
            items = self._seq("}")
            if len(items) % 2:
                raise EDNError("odd map")
            out = {}
            for k, v in zip(items[::2], items[1::2]):
                out[_hashable(k)] = v
            return out
        if c == "[":
            self.i += 1
            return self._seq("]")
        if c == "(":
            self.i += 1
            return self._seq(")")
        if c == '"':
            return self._string()
        if c == "\\":
            return self._char()
        if c == "#":
            return self._dispatch()
        if c in ")]}":
            raise EDNError(f"unexpected {c} at {self.i}")
        if c == "^":  # metadata: skip the meta form, return the value
            self.i += 1
            self.read()
            return self.read()
        if c == "'":
            self.i += 1
            return self.read()
        return self._atom()

    def _seq(self, close):
        out = []
        while True:
            self.skip()
            if self.i >= self.n:
                raise EDNError("eof in seq")
            if self.s[self.i] == close:
                self.i += 1
                return out
            out.append(self.read())

    def _string(self):
        self.i += 1
        buf = []
        s = self.s
        while True:
            if self.i >= self.n:
                raise EDNError("eof in string")
            c = s[self.i]
            if c == '"':
                self.i += 1
                return "".join(buf)
            if c == "\\":
                self.i += 1
                e = s[self.i]
                buf.append({"n": "\n", "t": "\t", "r": "\r"}.get(e, e))
            else:
                buf.append(c)
            self.i += 1

    def _char(self):
        self.i += 1
        j = self.i + 1
        while j < self.n and self.s[j] not in _DELIM:
            j += 1
        tok = self.s[self.i:j]
        self.i = j
        return {"newline": "\n", "space": " ", "tab": "\t"}.get(tok, tok[:1])

    def _dispatch(self):
        s = self.s
        nxt = s[self.i + 1] if self.i + 1 < self.n else ""
        if nxt == "{":
            self.i += 2
            return [x for x in self._seq("}")]  # set -> list
        if nxt == "_":
            self.i += 2
            self.read()
            return self.read()
        if nxt == "'":
            self.i += 2
            return self.read()
        # tagged literal such as #object[...] or #inst "..."
        self.i += 1
        j = self.i
        while j < self.n and s[j] not in _DELIM:
            j += 1
        tag = s[self.i:j]
        self.i = j
        return Tagged(tag, self.read())

    def _atom(self):
        j = self.i
        s = self.s
        while j < self.n and s[j] not in _DELIM:
            j += 1
        tok = s[self.i:j]
        self.i = j
        if not tok:
            raise EDNError("empty token")
        if tok == "nil":
            return None
        if tok == "true":
            return True
        if tok == "false":
            return False
        if tok[0] == ":":
            return tok
        try:
            if tok.endswith("N") or tok.endswith("M"):
                tok2 = tok[:-1]
            else:
                tok2 = tok
            if "/" in tok2 and tok2.replace("/", "").lstrip("+-").isdigit():
                f = Fraction(tok2)
                return float(f)
            if any(ch in tok2 for ch in ".eE") and tok2.lstrip("+-")[:1].isdigit():
                return float(tok2)
            return int(tok2)
        except ValueError:
            pass
        if tok in ("##Inf", "Infinity"):
            return float("inf")
        if tok in ("##NaN", "NaN"):
            return float("nan")
        return Sym(tok)


def _hashable(k):
    if isinstance(k, list):
        return tuple(_hashable(x) for x in k)
    if isinstance(k, dict):
        return tuple(sorted((kk, _hashable(v)) for kk, v in k.items()))
    if isinstance(k, Tagged):
        return repr(k)
    return k


def loads(text: str):
    return _Reader(text).read()


def read_from(text: str, pos: int):
    """Read one form starting at `pos`; returns (value, end_pos)."""
    r = _Reader(text, pos)
    v = r.read()
    return v, r.i
