"""Make an UNMODIFIED (possibly read-only) NPBench checkout see the b200 plugin.

NPBench resolves framework_info/ and the benchmark modules relative to the location of
npbench/infrastructure/framework.py (framework.py:22-24, 59-62; benchmark.py:20-22) and
only works from a source tree (setup.py:17 does not package npbench.benchmarks).  When the
checkout cannot be written to, `build_overlay` creates a directory that mirrors the
reference's npbench/, bench_info/ and framework_info/ with symlinks (Python keeps the
symlink path in __file__, so the harness looks inside the overlay), adds the plugin files
of npbench_b200/plugin/, and writes npbench/infrastructure/__init__.py as the reference's
text plus the single registration line frameworks.md:39-42 asks for.
`run_cli` then executes the reference's own run_benchmark.py / run_framework.py byte for
byte via runpy (a symlinked script would put the reference dir back on sys.path[0]).
"""
import importlib.util
import os
import runpy
import shutil
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
PLUGIN = os.path.join(_HERE, "plugin")
REGISTRATION_LINE = "from .b200_framework import *\n"


def _mirror(src_root: str, dst_root: str) -> None:
    for dirpath, dirnames, filenames in os.walk(src_root):
        dirnames[:] = [d for d in dirnames if d != "__pycache__"]
        rel = os.path.relpath(dirpath, src_root)
        out = os.path.normpath(os.path.join(dst_root, rel))
        os.makedirs(out, exist_ok=True)
        for f in filenames:
            dst = os.path.join(out, f)
            if not os.path.lexists(dst):
                os.symlink(os.path.join(dirpath, f), dst)


def _install_plugin(dst_root: str) -> None:
    for sub in ("framework_info", "npbench"):
        for dirpath, _, filenames in os.walk(os.path.join(PLUGIN, sub)):
            rel = os.path.relpath(dirpath, PLUGIN)
            out = os.path.join(dst_root, rel)
            os.makedirs(out, exist_ok=True)
            for f in filenames:
                if f.endswith((".py", ".json")):
                    dst = os.path.join(out, f)
                    if os.path.lexists(dst):
                        os.remove(dst)
                    shutil.copyfile(os.path.join(dirpath, f), dst)


def build_overlay(reference: str, dest: str) -> str:
    reference = os.path.abspath(reference)
    dest = os.path.abspath(dest)
    for sub in ("npbench", "bench_info", "framework_info"):
        if not os.path.isdir(os.path.join(reference, sub)):
            raise FileNotFoundError("%s is not an NPBench checkout (missing %s/)" % (reference, sub))
        _mirror(os.path.join(reference, sub), os.path.join(dest, sub))
    _install_plugin(dest)
    init = os.path.join(dest, "npbench", "infrastructure", "__init__.py")
    text = open(os.path.join(reference, "npbench", "infrastructure", "__init__.py")).read()
    if os.path.lexists(init):
        os.remove(init)
    with open(init, "w") as f:
        f.write(text if text.endswith("\n") else text + "\n")
        if REGISTRATION_LINE not in text:
            f.write(REGISTRATION_LINE)
    return dest


def prepare_sys_path(overlay: str) -> None:
    repo_root = os.path.dirname(_HERE)
    paths = [overlay, repo_root]
    if importlib.util.find_spec("pygount") is None:
        paths.append(os.path.join(PLUGIN, "pygount_stub"))
    for p in reversed(paths):
        if p in sys.path:
            sys.path.remove(p)
        sys.path.insert(0, p)


def run_cli(reference: str, overlay: str, script: str, argv) -> None:
    """Execute <reference>/<script> (e.g. run_benchmark.py) unmodified with `argv`."""
    build_overlay(reference, overlay)
    prepare_sys_path(overlay)
    old = sys.argv
    sys.argv = [os.path.join(reference, script)] + list(argv)
    try:
        runpy.run_path(os.path.join(reference, script), run_name="__main__")
    finally:
        sys.argv = old
