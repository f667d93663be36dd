"""N-rank == 1-rank check of the sharded evaluation: every rank evaluates its LPT share of ONE video list,
one all_gather brings the scores together, and the list-order score vector must equal the single-GPU vector BIT FOR BIT
(each video is computed by exactly one rank with kernels whose results do not depend on the batch they run in).

    python scripts/mgpu_determinism.py [--videos 160] [--workload xd]                       # 1 rank
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 \
        scripts/mgpu_determinism.py [--videos 160]                                            # 2 ranks
Rank 0 prints one JSON line: {"world", "sha256", "AUC", "AP", "frames"}."""
import argparse
import hashlib
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--videos", type=int, default=160)
    ap.add_argument("--workload", default="xd", choices=["ucf", "xd"])
    ap.add_argument("--host-inputs", action="store_true")
    args = ap.parse_args()
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    from iefvad_b200 import synth
    from iefvad_b200.evaluate import Evaluator
    from iefvad_b200.imf_vad import MMFMIL
    model = synth.build_model(MMFMIL, seed=0).to(dev).eval()
    T = synth.config_lengths(args.workload)[:args.videos]
    classes = synth.config_classes(args.workload, len(T))
    ev = Evaluator(model, T, classes, synth.make_gt(T, classes), rank=rank, world=world, device=dev)
    fi, fe = [], []
    for v in ev.mine:
        a, b = synth.make_video(int(v), int(T[v]))
        fi.append(a)
        fe.append(b)
    with torch.no_grad():
        if args.host_inputs:
            ev.set_host_ragged(fi, fe)
            res = ev.step(host_inputs=True)
        else:
            ev.set_device_features(ev.chunk_features(fi), ev.chunk_features(fe))
            res = ev.step()
    scores = res["scores"].cpu().numpy()
    if rank == 0:
        print(json.dumps({"world": world, "sha256": hashlib.sha256(scores.tobytes()).hexdigest(), "AUC": res["AUC"],
                          "AP": res["AP"], "frames": int(scores.size), "ano_AUC": res["ano_AUC"]}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
