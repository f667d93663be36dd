"""CPU tests of the host-side logic: torch-CPU port vs the numpy oracle, chunking rules, video partitioning and the
multi-rank gather layout (world_size 2 over gloo)."""
import os
import socket

import numpy as np
import pytest
import torch

from conftest import load_golden, params_from_npz
from oracle import iefvad_oracle as O
from oracle import torch_port

from iefvad_b200 import synth
from iefvad_b200.evaluate import gather_layout, num_chunks, partition_videos


@pytest.mark.parametrize("name", ["small_studentt", "small_gaussian", "small_r0"])
def test_torch_port_matches_reference_golden(name):
    z = load_golden(name + ".npz")
    P = {k: torch.from_numpy(v) for k, v in params_from_npz(z).items()}
    out = torch_port.forward(P, torch.from_numpy(z["img"]), torch.from_numpy(z["ev"]), heads=int(z["heads"]),
                             lambda_ref=float(z["lambda_ref"]), noise_model=str(z["noise_model"]), nu=float(z["nu"]))
    for k in out:
        assert O.max_norm_err(out[k].numpy(), z["out:" + k]) < 2e-5, k


def test_chunk_video_equals_reference_rule():
    for T in (1, 15, 255, 256, 257, 511, 512, 700):
        f = torch.arange(T * 4, dtype=torch.float32).reshape(T, 4) + 1
        ours = synth.chunk_video(f).numpy()
        ref, n = O.process_split(f.numpy(), 256)
        if ref.ndim == 2:
            ref = ref[None]
        assert n == T and np.array_equal(ours, ref)
        assert ours.shape[0] == num_chunks(T)


def test_partition_is_deterministic_balanced_and_complete():
    lengths = synth.config_lengths("xd")
    for world in (1, 2, 4, 8):
        parts = partition_videos(lengths, world)
        assert sorted(v for p in parts for v in p) == list(range(len(lengths)))
        loads = [sum(num_chunks(int(lengths[v])) for v in p) for p in parts]
        assert max(loads) - min(loads) <= max(num_chunks(int(t)) for t in lengths)
        assert parts == partition_videos(lengths, world)


def test_config_shapes():
    T = synth.config_lengths("ucf")
    assert len(T) == 290 and T.max() <= 4096 and list(T[:8]) == [1, 15, 16, 255, 256, 257, 512, 4096]
    classes = synth.config_classes("ucf", 290)
    assert set(classes) == set(synth.UCF_CLASSES) and classes.count("Normal") == 145
    gt = synth.make_gt(T, classes)
    assert gt.size == 16 * T.sum()
    assert len(synth.config_lengths("xd")) == 800


def _worker(rank, world, port, lengths, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        parts = partition_videos(lengths, world)
        src, dst, ln, counts, max_count = gather_layout(lengths, parts)
        goff = np.concatenate([[0], np.cumsum(lengths)])[:-1]
        # this rank's "scores": value encodes (video, row) so any misplacement is visible
        packed = torch.full((max(max_count, 1),), -1.0, dtype=torch.float64)
        off = 0
        for v in parts[rank]:
            packed[off:off + lengths[v]] = torch.arange(lengths[v], dtype=torch.float64) + 1e6 * v
            off += int(lengths[v])
        gathered = [torch.empty_like(packed) for _ in range(world)]
        dist.all_gather(gathered, packed)
        allv = torch.cat(gathered).numpy()
        out = np.full(int(lengths.sum()), np.nan)
        for s, d, n in zip(src, dst, ln):
            out[d:d + n] = allv[s:s + n]
        expect = np.concatenate([np.arange(lengths[v]) + 1e6 * v for v in range(len(lengths))])
        q.put((rank, bool(np.array_equal(out, expect)), counts))
    finally:
        dist.destroy_process_group()


def test_two_rank_gather_layout_over_gloo():
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    lengths = np.array([5, 300, 1, 256, 77, 1024, 33, 2, 600], dtype=np.int64)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, lengths, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok, _ in res)
    assert sum(res[0][2]) == lengths.sum()
