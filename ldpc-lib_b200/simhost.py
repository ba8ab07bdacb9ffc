"""Multi-process (one rank per GPU) mirror of the reference's bp_simulation() frame loop
(bp_simulation.cpp:591-824, 840) on top of Decoder.simulate().

Frames are independent, so rank g of G decodes the stripe [base + g*B, base + (g+1)*B) of every round and the
ranks exchange only error counters: one all-reduce (sum) of 4 integers per round.  The reference's stop rules
(`nde < n_frame_errors && experiment <= n_experiments`, and the early abort `nde >= 10 && nde / experiment >
2.5 * reference_frame_error`) are defined on the SEQUENTIAL frame order; to cut at exactly the frame where the
reference would, the per-frame records of a round are all-gathered and scanned in order -- but only for the
rounds in which a rule can possibly fire (decided from the counters alone), i.e. normally just the last one.
The result therefore does not depend on the number of ranks or on the round size.

Works with any torch.distributed backend (nccl on GPUs, gloo in the CPU tests) or without torch at all when
world_size == 1.  `simulate(first_frame, n_frames) -> uint32 records` is injectable for tests; records follow
ldpcb200_simulate(): bit 31 = frame in error, bit 30 = decoder reported success, bits 0..23 = info-bit errors.
"""
import numpy as np


class FrameLoopResult(dict):
    __getattr__ = dict.__getitem__


def _rank_world(group):
    if group is None:
        return 0, 1
    import torch.distributed as dist
    return dist.get_rank(group), dist.get_world_size(group)


def frame_loop(simulate, n_frame_errors, n_experiments, reference_frame_error, n_info_bits, round_frames=1 << 14,
               max_round_frames=1 << 18, group=None, device=None, first_frame=0):
    """Returns FrameLoopResult(ber, fer, experiment, nde, nse, nue, decoded, rounds, gathers)."""
    rank, world = _rank_world(group)
    if world > 1:
        import torch
        import torch.distributed as dist
    nde = nse = nue = experiment = decoded = rounds = gathers = 0
    limit = n_experiments + 1                       # `experiment <= n_experiments` is tested before the increment
    base = first_frame
    per_rank = round_frames
    stop = False
    while not stop and nde < n_frame_errors and experiment < limit:
        want = min(limit - experiment, per_rank * world)
        cnt = [want // world + (1 if g < want % world else 0) for g in range(world)]
        off = np.concatenate([[0], np.cumsum(cnt)])
        rec = simulate(base + int(off[rank]), cnt[rank]) if cnt[rank] else np.zeros(0, np.uint32)
        rec = np.ascontiguousarray(rec, dtype=np.uint32)
        err = (rec >> 31).astype(bool)
        local = np.array([cnt[rank], int(err.sum()), int((rec[err] & 0xFFFFFF).sum()), int(((rec[err] >> 30) & 1).sum())], np.int64)
        if world > 1:
            t = torch.from_numpy(local.copy())
            if device is not None:
                t = t.to(device)
            dist.all_reduce(t, group=group)          # the path's one collective: a few integers per round
            tot = t.cpu().numpy()
        else:
            tot = local
        rounds += 1
        decoded += want
        r_frames, r_nde, r_nse, r_nue = (int(x) for x in tot)
        # can a stop rule fire strictly inside this round?
        may_stop = nde + r_nde >= n_frame_errors
        if nde + r_nde >= 10 and r_nde > 0:
            bound = (nde + r_nde) / (experiment + r_nde)          # largest nde / experiment any prefix can reach
            may_stop |= bound > 2.5 * reference_frame_error
        if not may_stop:
            experiment += r_frames; nde += r_nde; nse += r_nse; nue += r_nue
        else:
            gathers += 1
            if world > 1:
                m = max(cnt)
                pad = np.zeros(m, np.uint32); pad[:cnt[rank]] = rec
                mine = torch.from_numpy(pad.view(np.int32).copy())
                if device is not None:
                    mine = mine.to(device)
                parts = [torch.empty_like(mine) for _ in range(world)]
                dist.all_gather(parts, mine, group=group)
                allrec = np.concatenate([p.cpu().numpy().view(np.uint32)[:cnt[g]] for g, p in enumerate(parts)])
            else:
                allrec = rec
            for w in allrec:                          # the reference's loop body after the decoder call, in frame order
                if not (nde < n_frame_errors and experiment < limit):
                    stop = True
                    break
                experiment += 1
                if w >> 31:
                    nse += int(w & 0xFFFFFF)
                    nde += 1
                    nue += int((w >> 30) & 1)
                    if nde >= 10 and nde / experiment > 2.5 * reference_frame_error:
                        stop = True
                        break
        base += want
        per_rank = min(per_rank * 4, max_round_frames)
    ber = nse / experiment / n_info_bits if experiment else 0.0
    fer = nde / experiment if experiment else 0.0
    return FrameLoopResult(ber=ber, fer=fer, experiment=experiment, nde=nde, nse=nse, nue=nue, decoded=decoded,
                           rounds=rounds, gathers=gathers)


def bp_simulation(dec, max_iterations, n_frame_errors, n_experiments, snr_db, reference_frame_error, modulation=0,
                  punctured_blocks=0, seed=1, stream=0, group=None, device=None, **kw):
    """bp_simulation(...) -> (BER, FER) for an open pyldpcb200.Decoder (one per rank), sharded over `group`."""
    def simulate(first, n):
        return dec.simulate(snr_db, n, max_iterations, modulation=modulation, punct=punctured_blocks, seed=seed, stream=stream,
                            first_frame=first, want_per_frame=True)["per_frame"]
    r = frame_loop(simulate, n_frame_errors, n_experiments, reference_frame_error, dec.K, group=group, device=device, **kw)
    return r.ber, r.fer, r
