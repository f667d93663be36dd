"""Recipe for `oracle/_ref/`: a byte-for-byte snapshot of the reference's own source files for this path.

TEST / BASELINE INFRASTRUCTURE ONLY.  EavnJeong/IEF-VAD is pure Python, so "building the reference" means placing its
unmodified modules where the GPU box can import them (`/root/reference` does not exist there).  `oracle/_ref/` is
git-ignored (the reference's sources never enter this repository's history) but not gpurun-ignored, so it travels
with the snapshot like the built `libiefvad.so`.

    python oracle/make_ref.py            # copies from $IEFVAD_REFERENCE (default /root/reference)

Files (all imported as-is by `bench.py --impl reference[-cuda]` and by nothing in the product package):
    model/imf_vad.py   MMFMIL / MultiModal_Fusion_Attn_Iter   (the forward under test)
    data/tools.py      process_split                          (the caller's chunk / zero-pad rule)
    train/loss.py      CLAS2                                  (config 4)
    train/ucf_test.py  test()  (+ train/utils.py, train/metrics.py it imports)   the reference's evaluation loop, run
                       UNMODIFIED against the B200 module by tests/test_gpu_dropin.py
A `MANIFEST.json` with the sha256 of every copied file is written beside them; `verify()` re-checks it so a
tampered copy is detected before it is timed."""
from __future__ import annotations

import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")
FILES = ("model/imf_vad.py", "data/tools.py", "train/loss.py",
         # the reference's own evaluation loop (the CALLER of the path), for the drop-in test that runs it unmodified against
         # the B200 module (tests/test_gpu_dropin.py); it imports .utils and .metrics relatively
         "train/ucf_test.py", "train/utils.py", "train/metrics.py")


def _sha(path: str) -> str:
    with open(path, "rb") as fh:
        return hashlib.sha256(fh.read()).hexdigest()


def make(reference: str | None = None) -> str | None:
    """Copy FILES from the reference tree into oracle/_ref/.  Returns the destination, or None when the reference
    tree is not present (GPU box: the prebuilt copy is used)."""
    reference = reference or os.environ.get("IEFVAD_REFERENCE", "/root/reference")
    if not os.path.isdir(reference):
        return None
    manifest = {}
    for rel in FILES:
        src = os.path.join(reference, rel)
        dst = os.path.join(DEST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        manifest[rel] = _sha(dst)
    with open(os.path.join(DEST, "MANIFEST.json"), "w") as fh:
        json.dump({"source": reference, "files": manifest}, fh, indent=1)
    return DEST


def available() -> bool:
    return os.path.exists(os.path.join(DEST, "MANIFEST.json"))


def verify() -> bool:
    try:
        with open(os.path.join(DEST, "MANIFEST.json")) as fh:
            files = json.load(fh)["files"]
        return all(_sha(os.path.join(DEST, rel)) == h for rel, h in files.items()) and set(files) == set(FILES)
    except Exception:
        return False


def import_reference():
    """-> (MMFMIL, process_split, CLAS2) imported from oracle/_ref (unmodified reference modules).  The modules are
    loaded under private names so that they never shadow or get shadowed by a `model` / `data` / `train` package."""
    import importlib.util
    if not verify():
        raise ImportError("oracle/_ref is missing or does not match its manifest; run `python oracle/make_ref.py` "
                          "where /root/reference exists")
    mods = []
    for rel in FILES[:3]:
        name = "_iefvad_ref_" + rel[:-3].replace("/", "_")
        spec = importlib.util.spec_from_file_location(name, os.path.join(DEST, rel))
        mod = importlib.util.module_from_spec(spec)
        sys.dont_write_bytecode = True
        spec.loader.exec_module(mod)
        mods.append(mod)
    return mods[0].MMFMIL, mods[1].process_split, mods[2].CLAS2


def import_reference_eval_loop():
    """-> the reference's `train.ucf_test` module (its `test()` is the per-video evaluation loop, train/ucf_test.py:16-178),
    imported from oracle/_ref as the package `train` with no-op stubs for the plotting / logging packages it imports
    (matplotlib, wandb - neither is installed, neither touches the numbers)."""
    import importlib
    import types
    if not verify():
        raise ImportError("oracle/_ref is missing or does not match its manifest; run `python oracle/make_ref.py` "
                          "where /root/reference exists")
    for name in ("matplotlib", "matplotlib.pyplot", "wandb"):
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.log = lambda *a, **k: None
            m.init = lambda *a, **k: None
            sys.modules[name] = m
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.dont_write_bytecode = True
    if DEST not in sys.path:
        sys.path.insert(0, DEST)
    for k in [k for k in sys.modules if k == "train" or k.startswith("train.")]:
        del sys.modules[k]
    return importlib.import_module("train.ucf_test")


if __name__ == "__main__":
    d = make()
    print(d if d else "reference tree not found; nothing copied")
