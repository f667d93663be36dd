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
FILES = ("model/imf_vad.py", "data/tools.py", "train/loss.py")


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
    for rel in FILES:
        name = "_iefvad_ref_" + rel[:-3].replace("/", "_")
        spec = importlib.util.spec_from_file_location(name, os.path.join(DEST, rel))
        mod = importlib.util.module_from_spec(spec)
        sys.dont_write_bytecode = True
        spec.loader.exec_module(mod)
        mods.append(mod)
    return mods[0].MMFMIL, mods[1].process_split, mods[2].CLAS2


if __name__ == "__main__":
    d = make()
    print(d if d else "reference tree not found; nothing copied")
