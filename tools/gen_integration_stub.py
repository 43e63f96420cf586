"""Regenerate the ctypes stub of INTEGRATION.md section 2 from redback_b200/_capi.py (the structures a foreign caller has
to mirror), between the `<!-- stub:begin -->` / `<!-- stub:end -->` markers.  tests/test_capi_symbols.py checks that
the committed text equals this output, so the document cannot drift from the header again.

    python tools/gen_integration_stub.py            # print
    python tools/gen_integration_stub.py --write    # rewrite INTEGRATION.md in place
"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _ctype_name(t):
    if isinstance(t, type) and issubclass(t, C.Structure):
        return t.__name__
    if isinstance(t, type) and issubclass(t, C.Array):
        return f"C.{t._type_.__name__} * {t._length_}"
    return "C." + t.__name__


def struct_source(cls, comment=""):
    items = [f'("{n}", {_ctype_name(t)})' for n, t in cls._fields_]
    lines, cur = [], "    _fields_ = ["
    for i, item in enumerate(items):
        piece = item + ("," if i + 1 < len(items) else "]")
        if len(cur) + 1 + len(piece) > 112 and not cur.endswith("["):
            lines.append(cur.rstrip())
            cur = " " * 16 + piece
        else:
            cur += ("" if cur.endswith("[") else " ") + piece
    lines.append(cur)
    return f"class {cls.__name__}(C.Structure):{comment}\n" + "\n".join(lines) + "\n"


def stub():
    from redback_b200 import _capi
    parts = ["```python", "import ctypes as C",
             'lib = C.CDLL("redback_b200/_lib/libredback_b200.so")', "lib.rb_sizeof.restype = C.c_int64", "",
             struct_source(_capi.Cosmology), struct_source(_capi.UpperLimits),
             struct_source(_capi.SnConfig, "          # rb_sn_config, field for field (include/redback_b200.h)"),
             "assert lib.rb_sizeof(2) == C.sizeof(SnConfig)      # 2 = rb_sn_config: layout check before the first call",
             "cfg = SnConfig(); lib.rb_sn_config_default(C.byref(cfg))",
             "cfg.engine, cfg.sed = 1, 2           # RB_ENGINE_MAGNETAR, RB_SED_CUTOFF_BLACKBODY  == redback's `slsn`",
             "# host buffers: params is float64 [RB_SN_NPARAM=10][N] (row order in the header), t/nu/y/sigma are [E]",
             "rc = lib.rb_sn_eval_f64_host(C.byref(cfg), N, E, params.ctypes.data_as(C.c_void_p), t.ctypes.data_as(C.c_void_p),",
             "                             nu.ctypes.data_as(C.c_void_p), None, y.ctypes.data_as(C.c_void_p),",
             "                             sigma.ctypes.data_as(C.c_void_p), None, None, logl.ctypes.data_as(C.c_void_p))",
             "assert rc == 0, lib.rb_last_error()", "```"]
    return "\n".join(parts) + "\n"


BEGIN, END = "<!-- stub:begin -->\n", "<!-- stub:end -->\n"


def main():
    text = stub()
    if "--write" in sys.argv:
        path = os.path.join(ROOT, "INTEGRATION.md")
        doc = open(path).read()
        a, b = doc.index(BEGIN) + len(BEGIN), doc.index(END)
        open(path, "w").write(doc[:a] + text + doc[b:])
    else:
        sys.stdout.write(text)


if __name__ == "__main__":
    main()
