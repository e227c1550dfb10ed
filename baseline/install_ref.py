"""Install the unmodified reference (wilson-labs/cola, read-only at /root/reference in the build container) into
baseline/_ref so that it travels to the GPU box:

    python -m pip install --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse \
           --target baseline/_ref <copy of /root/reference under /tmp>

(`--no-deps`: the reference's pinned `cola-plum-dispatch==0.1.4` and `optree` are not in the offline wheelhouse; a copy
under /tmp because the build writes egg-info into the source tree and /root/reference is read-only.)  The two missing
packages are replaced by this repo's own import shims (tests/golden/refshim/{plum,optree}, written for the golden
generator: with them the reference's non-JAX test-suite gives 290 passed / 1 skipped), copied next to the package.
Nothing of the reference is committed: baseline/_ref/ is git-ignored.

    python baseline/install_ref.py        # (re)install; prints the path

`reference_sys_path()` is what tests and bench.py put on sys.path to import `cola` = the reference: the live
/root/reference tree where it exists (build container), else the installed copy (GPU box)."""
import json
import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF_SRC = "/root/reference"
TARGET = os.path.join(HERE, "_ref")
SHIMS = os.path.join(ROOT, "tests", "golden", "refshim")


def installed():
    return os.path.isdir(os.path.join(TARGET, "cola")) and os.path.isdir(os.path.join(TARGET, "plum"))


def install(force=False, verbose=False):
    """Returns the install path, or None when neither the reference tree nor an earlier install is available."""
    if not os.path.isdir(os.path.join(REF_SRC, "cola")):
        return TARGET if installed() else None
    if installed() and not force:
        return TARGET
    shutil.rmtree(TARGET, ignore_errors=True)
    os.makedirs(TARGET, exist_ok=True)
    with tempfile.TemporaryDirectory(prefix="cola_ref_") as tmp:
        src = os.path.join(tmp, "reference")
        shutil.copytree(REF_SRC, src, ignore=shutil.ignore_patterns(".git", "docs", "*.ipynb"))
        cmd = [sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--no-deps", "--find-links",
               "/opt/wheelhouse", "--target", TARGET, src]
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if verbose:
            print(r.stdout)
        if r.returncode != 0:
            raise RuntimeError("pip install of the reference failed:\n" + r.stdout[-2000:])
    # The reference's setup.py uses find_packages(), which skips its directories without an __init__.py
    # (cola/linalg/tbd, svd, preconditioning, ... are implicit namespace packages in the source tree): complete the
    # installed package with every source file pip left out, unmodified.
    missing = []
    for dirpath, dirnames, filenames in os.walk(os.path.join(REF_SRC, "cola")):
        dirnames[:] = [d for d in dirnames if d != "__pycache__"]
        rel = os.path.relpath(dirpath, REF_SRC)
        for fn in filenames:
            if fn.endswith(".py") and not os.path.exists(os.path.join(TARGET, rel, fn)):
                os.makedirs(os.path.join(TARGET, rel), exist_ok=True)
                shutil.copy2(os.path.join(dirpath, fn), os.path.join(TARGET, rel, fn))
                missing.append(os.path.join(rel, fn))
    for shim in ("plum", "optree"):
        shutil.copytree(os.path.join(SHIMS, shim), os.path.join(TARGET, shim),
                        ignore=shutil.ignore_patterns("__pycache__"))
    with open(os.path.join(TARGET, "INSTALL.json"), "w") as f:
        json.dump({"source": REF_SRC, "how": "pip install --no-index --no-build-isolation --no-deps --target baseline/_ref",
                   "shims": ["plum", "optree"], "completed_namespace_files": missing,
                   "note": "unmodified reference package + this repo's import shims"}, f)
    return TARGET


def reference_sys_path():
    """sys.path entries (in order) that make `import cola` resolve to the reference, or [] when it is unavailable."""
    if os.path.isdir(os.path.join(REF_SRC, "cola")) and not os.environ.get("COLA_REF_FORCE_INSTALLED"):
        return [SHIMS, REF_SRC]
    if installed():
        return [TARGET]
    return []


def import_reference():
    """`import cola` = the reference (raises ImportError when unavailable).  Byte-code is not written next to it."""
    paths = reference_sys_path()
    if not paths:
        raise ImportError("the reference is neither at /root/reference nor installed in baseline/_ref "
                          "(run `python baseline/install_ref.py` in the build container)")
    sys.dont_write_bytecode = True
    for p in reversed(paths):
        if p not in sys.path:
            sys.path.insert(0, p)
    import cola
    assert any(os.path.abspath(cola.__file__).startswith(os.path.abspath(p)) for p in paths), cola.__file__
    return cola


if __name__ == "__main__":
    print(install(force=True, verbose="-v" in sys.argv))
