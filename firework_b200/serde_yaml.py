"""serde_yaml 0.8-style emitter for firework `Scene` documents (block style, 2-space indent, sequences
indented under their key, f32 values printed as the shortest f64 text that round-trips, `~` for None).

Matches the layout of the reference's committed scenes (scenes/conics.yml) so that documents written here
are readable by `serde_yaml::from_reader::<Scene>` and by the native loader alike.  `loads` (PyYAML) is
used by the test harness only; the product parses YAML in C++ (csrc/yaml_lite.cpp).
"""
from __future__ import annotations

import io


def _scalar(v) -> str:
    if v is None:
        return "~"
    if v is True:
        return "true"
    if v is False:
        return "false"
    if isinstance(v, int):
        return str(v)
    if isinstance(v, float):
        if v != v:
            return ".nan"
        if v in (float("inf"), float("-inf")):
            return ".inf" if v > 0 else "-.inf"
        r = repr(v)
        if "e" in r or "E" in r:
            # serde_yaml prints exponents as e.g. 1e-5; keep a form every YAML 1.1 float regex accepts
            m, e = r.split("e")
            if "." not in m:
                m += ".0"
            return f"{m}e{int(e):+d}"
        if "." not in r:
            r += ".0"
        return r
    s = str(v)
    if s == "" or s[0] in "-?:,[]{}#&*!|>'\"%@`~" or ": " in s or " #" in s or s in ("true", "false", "null", "~"):
        return '"' + s.replace("\\", "\\\\").replace('"', '\\"') + '"'
    return s


def _emit(out: io.StringIO, v, indent: int, inline_first: bool):
    """Emit mapping/sequence `v`. `inline_first` = the first line continues after a '- '."""
    pad = " " * indent
    if isinstance(v, dict):
        first = True
        for k, val in v.items():
            prefix = "" if (first and inline_first) else pad
            first = False
            if isinstance(val, dict) and val:
                out.write(f"{prefix}{k}:\n")
                _emit(out, val, indent + 2, False)
            elif isinstance(val, list) and val:
                out.write(f"{prefix}{k}:\n")
                _emit(out, val, indent + 2, False)
            elif isinstance(val, (dict, list)):
                out.write(f"{prefix}{k}: {'{}' if isinstance(val, dict) else '[]'}\n")
            else:
                out.write(f"{prefix}{k}: {_scalar(val)}\n")
    elif isinstance(v, list):
        for item in v:
            if isinstance(item, (dict, list)) and item:
                out.write(f"{pad}- ")
                _emit(out, item, indent + 2, True)
            else:
                out.write(f"{pad}- {_scalar(item)}\n")
    else:
        raise TypeError(type(v))


def dumps(doc: dict) -> str:
    out = io.StringIO()
    out.write("---\n")
    _emit(out, doc, 0, False)
    return out.getvalue()


def loads(text: str) -> dict:
    import yaml
    return yaml.load(text, Loader=getattr(yaml, "CSafeLoader", yaml.SafeLoader))
