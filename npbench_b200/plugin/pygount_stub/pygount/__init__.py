"""Minimal stand-in for `pygount`, used ONLY when the real package is not installed.

npbench/infrastructure/line_count.py:6 imports pygount at module scope, so the harness
cannot even be imported without it.  LineCount only needs
SourceAnalysis.from_file(path, group).code_count (line_count.py:17-88)."""


class SourceAnalysis:
    def __init__(self, code_count):
        self.code_count = code_count

    @classmethod
    def from_file(cls, path, group, *args, **kwargs):
        n = 0
        with open(path, errors="replace") as f:
            for line in f:
                s = line.strip()
                if s and not s.startswith("#"):
                    n += 1
        return cls(n)
