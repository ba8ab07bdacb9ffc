"""The sharded frame loop (ldpc-lib_b200/simhost.py): the reference's stop rules applied in frame order, results
independent of the number of ranks and of the round size.  Multi-rank runs use the gloo backend on CPU."""
import importlib.util
import os
import sys

import numpy as np
import pytest

from conftest import ROOT


def load_simhost():
    spec = importlib.util.spec_from_file_location("simhost", os.path.join(ROOT, "ldpc-lib_b200", "simhost.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def fake_records(first, n, p_err=0.03, seed=5):
    """Deterministic per-frame records keyed by the global frame index (what the counter-based generator gives)."""
    idx = np.arange(first, first + n, dtype=np.uint64)
    h = (idx * np.uint64(0x9E3779B97F4A7C15) + np.uint64(seed)) >> np.uint64(11)
    h = (h ^ (h >> np.uint64(17))) * np.uint64(0xD6E8FEB86659FD93) & np.uint64(0xFFFFFFFFFFFF)
    u = (h % np.uint64(1000003)).astype(np.float64) / 1000003.0
    err = u < p_err
    info = (1 + (h >> np.uint64(20)) % np.uint64(7)).astype(np.uint32)
    undet = ((h >> np.uint64(9)) & np.uint64(1)).astype(np.uint32)
    return np.where(err, np.uint32(0x80000000) | (undet << np.uint32(30)) | info, np.uint32(0)).astype(np.uint32)


def sequential_reference(n_frame_errors, n_experiments, ref_fer, **kw):
    """bp_simulation.cpp:591-824 literally, one frame at a time."""
    nde = nse = nue = experiment = 0
    while nde < n_frame_errors and experiment <= n_experiments:
        experiment += 1
        w = int(fake_records(experiment - 1, 1, **kw)[0])
        if w >> 31:
            nse += w & 0xFFFFFF
            nde += 1
            nue += (w >> 30) & 1
            if nde >= 10 and nde / experiment > 2.5 * ref_fer:
                break
    return experiment, nde, nse, nue


CASES = [dict(n_frame_errors=50, n_experiments=100000, ref_fer=1.0),        # stops on the error count
         dict(n_frame_errors=10 ** 6, n_experiments=7000, ref_fer=1.0),     # stops on the frame budget (runs n + 1 frames)
         dict(n_frame_errors=500, n_experiments=100000, ref_fer=0.01),      # early abort: FER > 2.5 x reference
         dict(n_frame_errors=12, n_experiments=30, ref_fer=1.0)]            # tiny


@pytest.mark.parametrize("case", CASES)
@pytest.mark.parametrize("round_frames", [7, 256, 1 << 14])
def test_single_rank_matches_sequential_loop(case, round_frames):
    sh = load_simhost()
    want = sequential_reference(case["n_frame_errors"], case["n_experiments"], case["ref_fer"])
    r = sh.frame_loop(lambda f, n: fake_records(f, n), case["n_frame_errors"], case["n_experiments"], case["ref_fer"], 4096,
                      round_frames=round_frames)
    assert (r.experiment, r.nde, r.nse, r.nue) == want
    assert r.fer == want[1] / want[0] and r.ber == want[2] / want[0] / 4096


def _worker(rank, world, port, case, round_frames, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sh = load_simhost()
    r = sh.frame_loop(lambda f, n: fake_records(f, n), case["n_frame_errors"], case["n_experiments"], case["ref_fer"], 4096,
                      round_frames=round_frames, group=dist.group.WORLD)
    q.put((rank, r.experiment, r.nde, r.nse, r.nue, r.gathers, r.rounds))
    dist.destroy_process_group()


@pytest.mark.parametrize("case_idx", [0, 1, 2])
def test_two_ranks_gloo_give_the_same_counts(case_idx):
    import torch.multiprocessing as mp
    case = CASES[case_idx]
    want = sequential_reference(case["n_frame_errors"], case["n_experiments"], case["ref_fer"])
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + case_idx + (os.getpid() % 200)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, case, 300, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for g in got:
        assert tuple(g[1:5]) == want, (g, want)
        assert g[5] <= g[6]                          # record gathers happen only in rounds where a rule can fire
