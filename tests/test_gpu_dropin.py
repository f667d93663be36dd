"""The drop-in claim, end to end: the reference's OWN evaluation loop (`train/ucf_test.py:test()`, imported unmodified from
oracle/_ref) is run against the B200 `MMFMIL` - the loop moves the module with `.to(device)`, calls `.eval()`, feeds it the
chunks `process_split` makes, slices `logits[:len]`, reads `w_i / w_e / fused / image_mu / event_mu` with `.cpu().numpy()`
and calls sklearn - and must return the AUC / AP the reference module produced in the same loop (tests/golden/eval_loop.npz,
made by running that loop with the reference module).  Needs a B200 and oracle/_ref (python oracle/make_ref.py)."""
import os
import types

import numpy as np
import pytest
import torch

from conftest import load_golden

pytestmark = pytest.mark.gpu


def test_reference_eval_loop_runs_unmodified_on_the_b200_module(tmp_path):
    from oracle import make_ref
    if not make_ref.available():
        pytest.skip("oracle/_ref is absent (python oracle/make_ref.py where /root/reference exists)")
    ref_ucf_test = make_ref.import_reference_eval_loop()
    _, process_split, _ = make_ref.import_reference()
    import iefvad_b200
    from iefvad_b200 import synth
    z = load_golden("eval_loop.npz")
    T, classes = z["lengths"], [str(c) for c in z["classes"]]
    gt = synth.make_gt(T, classes)
    model = synth.build_model(iefvad_b200.MMFMIL, seed=0)          # the loop itself calls model.to(device) and model.eval()

    class Loader:                                                    # what data/dataset.py:34-52 + DataLoader(batch_size=1) deliver
        def __iter__(self):
            for v in range(len(T)):
                img, ev = synth.make_video(100 + v, int(T[v]))
                fi, ln = process_split(img.numpy(), 256)
                fe, _ = process_split(ev.numpy(), 256)
                yield (torch.from_numpy(fi)[None], torch.from_numpy(fe)[None], [classes[v]], torch.tensor([ln]))

    captured = {}
    real_auc = ref_ucf_test.roc_auc_score

    def spy(gt_, pred):
        captured.setdefault("first", np.asarray(pred)[::16].copy())
        return real_auc(gt_, pred)

    cwd = os.getcwd()
    os.chdir(tmp_path)                                               # the loop creates ./vis
    try:
        ref_ucf_test.roc_auc_score = spy
        ret = ref_ucf_test.test(types.SimpleNamespace(exp_name="dropin", dataset="ucfcrime"), model, Loader(), 256, None, gt,
                                torch.device("cuda:0"))
    finally:
        ref_ucf_test.roc_auc_score = real_auc
        os.chdir(cwd)
    scores = captured["first"]
    assert scores.shape == z["scores"].shape
    err = float(np.max(np.abs(scores - z["scores"]) / z["scores"]))
    assert err < 1e-3, err                                           # per-frame scores, default plan (fp16 operands)
    assert abs(float(ret[0]) - float(z["AUC"])) < 2e-3 and abs(float(ret[1]) - float(z["AP"])) < 2e-3
