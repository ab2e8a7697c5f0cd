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
import hashlib
import importlib.util
import json
import os
import runpy
import shutil
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
_REPO = os.path.dirname(_HERE)
PLUGIN = os.path.join(_HERE, "plugin")
REGISTRATION_LINE = "from .b200_framework import *\n"
STAGED = os.path.join(_REPO, "baseline", "_ref")          # git-ignored; travels to the GPU box with the snapshot
_STAGE_ITEMS = ("npbench", "bench_info", "framework_info", "run_benchmark.py", "run_framework.py", "LICENSE")
MANIFEST = "B200_STAGE_MANIFEST.json"


def is_checkout(path: str) -> bool:
    return bool(path) and all(os.path.isdir(os.path.join(path, d)) for d in ("npbench", "bench_info", "framework_info"))


def find_reference(explicit: str = None) -> str:
    """An unmodified NPBench checkout: --reference / $NPBENCH_REF, /root/reference (build container),
    or the byte-for-byte staged copy under baseline/_ref (GPU box).  Raises when none is present."""
    tried = []
    for cand in (explicit, os.environ.get("NPBENCH_REF"), "/root/reference", STAGED):
        if cand:
            tried.append(cand)
            if is_checkout(cand):
                return os.path.abspath(cand)
    raise FileNotFoundError("no NPBench checkout found (tried %s); stage one with "
                            "`python -m npbench_b200.overlay --stage <checkout>`" % ", ".join(tried))


def _sha(path: str) -> str:
    with open(path, "rb") as f:
        return hashlib.sha256(f.read()).hexdigest()


def stage_reference(src: str, dest: str = STAGED) -> str:
    """Copy the harness-relevant part of an NPBench checkout, byte for byte, to baseline/_ref so that
    the UNMODIFIED harness can run on a box where /root/reference is not mounted.  Writes a manifest
    (relative path -> sha256 of the source file) that verify_staged() re-checks."""
    src, dest = os.path.abspath(src), os.path.abspath(dest)
    if not is_checkout(src):
        raise FileNotFoundError("%s is not an NPBench checkout" % src)
    if os.path.isdir(dest):
        shutil.rmtree(dest)
    os.makedirs(dest)
    manifest = {}
    for item in _STAGE_ITEMS:
        s = os.path.join(src, item)
        if os.path.isdir(s):
            for dirpath, dirnames, filenames in os.walk(s):
                dirnames[:] = sorted(d for d in dirnames if d != "__pycache__")
                for f in sorted(filenames):
                    if f.endswith((".pyc", ".db")):
                        continue
                    rel = os.path.relpath(os.path.join(dirpath, f), src)
                    os.makedirs(os.path.dirname(os.path.join(dest, rel)), exist_ok=True)
                    shutil.copyfile(os.path.join(src, rel), os.path.join(dest, rel))
                    manifest[rel] = _sha(os.path.join(src, rel))
        elif os.path.isfile(s):
            shutil.copyfile(s, os.path.join(dest, item))
            manifest[item] = _sha(s)
    with open(os.path.join(dest, MANIFEST), "w") as f:
        json.dump({"source": src, "files": manifest}, f, indent=0, sort_keys=True)
    return dest


def verify_staged(dest: str = STAGED):
    """[(relative path, problem)] for every staged file that no longer matches its manifest hash."""
    with open(os.path.join(dest, MANIFEST)) as f:
        files = json.load(f)["files"]
    bad = []
    for rel, h in files.items():
        p = os.path.join(dest, rel)
        if not os.path.isfile(p):
            bad.append((rel, "missing"))
        elif _sha(p) != h:
            bad.append((rel, "modified"))
    return bad


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
    reference = find_reference(reference)
    build_overlay(reference, overlay)
    prepare_sys_path(overlay)
    old = sys.argv
    sys.argv = [os.path.join(reference, script)] + list(argv)
    try:
        runpy.run_path(os.path.join(reference, script), run_name="__main__")
    finally:
        sys.argv = old


if __name__ == "__main__":      # python -m npbench_b200.overlay --stage <checkout>
    if len(sys.argv) == 3 and sys.argv[1] == "--stage":
        print(stage_reference(sys.argv[2]))
    else:
        print(find_reference())
