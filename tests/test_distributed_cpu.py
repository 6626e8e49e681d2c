"""N>1 host logic on CPU: two gloo processes on 127.0.0.1 run the data-parallel plumbing
(`ddpo_b200/training/distributed.py`) that wraps the one gradient all-reduce of the PPO update.

Checked against the oracle's AccumulatingTrainState (reference pipeline/policy_gradient.py:13-57, :137-142):
two ranks x two accumulated micro-gradients each, reduced with `allreduce_sum_` and scaled by `grad_scale`,
must give the same AdamW update on both ranks as ONE process that saw all four micro-gradients.
"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from ddpo_b200.training import distributed as D
from oracle import optim as O

N_PARAMS = 4099
N_ACC = 2
WORLD = 2


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _micro_grads():
    rng = np.random.default_rng(7)
    return rng.standard_normal((WORLD, N_ACC, N_PARAMS)).astype(np.float32) * 0.02


def _worker(rank, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(WORLD))
    dist.init_process_group("gloo", rank=rank, world_size=WORLD)
    try:
        assert D.is_distributed() and D.world_size() == WORLD and D.rank() == rank
        g = _micro_grads()
        # fused accumulation: the backward pass adds every micro-gradient into grad_acc
        grad_acc = torch.zeros(N_PARAMS)
        for i in range(N_ACC):
            grad_acc += torch.from_numpy(g[rank, i])
        world = D.allreduce_sum_(grad_acc)
        scale = D.grad_scale(N_ACC, world)
        mean_grad = (grad_acc * scale).numpy()
        # PPO info: lax.pmean over ranks
        info = torch.tensor([0.1 * (rank + 1), 0.5 * rank, -1.0 + rank], dtype=torch.float32)
        D.pmean_(info)
        lo, hi = D.shard_bounds(16)
        t = D.max_over_ranks(10.0 + rank)
        # optimizer update every rank applies to its replica
        params = np.linspace(-1, 1, N_PARAMS, dtype=np.float32)
        st = O.AdamWState(N_PARAMS)
        new_p, gn = O.clip_adamw_update(params, mean_grad, st, max_norm=1.0)
        np.savez(os.path.join(out_dir, f"rank{rank}.npz"), mean_grad=mean_grad, info=info.numpy(), lo=lo, hi=hi, t=t,
                 new_p=new_p, gn=gn, world=world)
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_two_rank_gradient_reduction_matches_single_process(tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(port, str(tmp_path)), nprocs=WORLD, join=True)
    r = [np.load(tmp_path / f"rank{i}.npz") for i in range(WORLD)]
    g = _micro_grads()
    # single process that accumulated all WORLD * N_ACC micro-gradients (reference :33-41)
    ref = O.AccumulatingTrainState(np.linspace(-1, 1, N_PARAMS, dtype=np.float32), max_norm=1.0)
    flat = g.reshape(-1, N_PARAMS)
    for i, gi in enumerate(flat):
        ref.apply_gradients(gi, do_update=(i == len(flat) - 1))
    for i in range(WORLD):
        assert int(r[i]["world"]) == WORLD
        np.testing.assert_allclose(r[i]["mean_grad"], flat.mean(0), rtol=0, atol=2e-8)
        np.testing.assert_allclose(r[i]["new_p"], ref.params, rtol=0, atol=2e-7)
        np.testing.assert_allclose(r[i]["info"], [0.15, 0.25, -0.5], rtol=0, atol=1e-7)
        assert (int(r[i]["lo"]), int(r[i]["hi"])) == (8 * i, 8 * i + 8)
        assert float(r[i]["t"]) == 11.0
    # replicas stay bit-identical: same reduced gradient -> same update on every rank
    assert np.array_equal(r[0]["mean_grad"], r[1]["mean_grad"])
    assert np.array_equal(r[0]["new_p"], r[1]["new_p"])


def test_single_process_helpers_are_identity():
    assert not D.is_distributed() and D.world_size() == 1 and D.rank() == 0
    t = torch.arange(5, dtype=torch.float32)
    assert D.allreduce_sum_(t) == 1 and torch.equal(t, torch.arange(5, dtype=torch.float32))
    assert torch.equal(D.pmean_(t.clone()), t)
    assert D.shard_bounds(8) == (0, 8)
    assert D.shard_bounds(8, 1, 4) == (2, 4)
    assert D.max_over_ranks(3.5) == 3.5
    assert D.grad_scale(4, 2) == 0.125
    with pytest.raises(ValueError):
        D.shard_bounds(7, 0, 2)
    with pytest.raises(ValueError):
        D.grad_scale(0, 1)


# ------------------------------------------------------------------ epoch driver: reward / advantage exchange ----
def _driver_worker(rank, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(WORLD))
    dist.init_process_group("gloo", rank=rank, world_size=WORLD)
    try:
        from ddpo_b200.pipeline import policy_gradient as PG
        from ddpo_b200.utils.stat_tracking import PerPromptStatTracker
        local_rewards = np.arange(4, dtype=np.float64)[:, None] * (rank + 1)        # jpeg-style [N, 1]
        local_prompts = np.array([f"p{(i + rank) % 2}" for i in range(4)])
        rewards = PG.allgather_array(local_rewards)                                  # process_allgather(tiled=True)
        prompts = PG.allgather_array(local_prompts)
        tracker = PerPromptStatTracker(32, 2)
        adv = PG.compute_advantages(rewards, prompts, tracker)
        mine = np.asarray(adv).reshape(WORLD, -1)[rank]                               # reference :349
        glob = PG.compute_advantages(rewards, prompts, None).reshape(WORLD, -1)[rank]
        # the driver keys the per-prompt statistics by tokenizer-decoded prompt ids gathered from every rank
        # (reference :329-335): a rank must decode prompts it never tokenized itself to the very same strings
        from ddpo_b200.utils.text_stub import StubTokenizer
        tok = StubTokenizer()
        own = [f"a {w} riding a bike" for w in (("zebra", "llama", "zebra", "yak") if rank == 0 else ("otter", "zebra", "emu", "emu"))]
        ids = PG.allgather_array(tok(own, padding="max_length", return_tensors="np").input_ids)
        decoded = np.array(tok.batch_decode(ids, skip_special_tokens=True))
        adv_tok = PerPromptStatTracker(32, 2).update(decoded, rewards)
        np.savez(os.path.join(out_dir, f"drv{rank}.npz"), rewards=rewards, prompts=prompts, mine=mine, glob=glob,
                 decoded=decoded, adv_tok=np.asarray(adv_tok))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_two_rank_reward_allgather_and_advantage_slices(tmp_path):
    """each worker sees the pod's rewards in rank order, normalises over ALL of them and keeps its own slice
    (reference pipeline/policy_gradient.py:323-349)"""
    port = _free_port()
    mp.spawn(_driver_worker, args=(port, str(tmp_path)), nprocs=WORLD, join=True)
    r = [np.load(tmp_path / f"drv{i}.npz") for i in range(WORLD)]
    want_rewards = np.concatenate([np.arange(4)[:, None] * 1.0, np.arange(4)[:, None] * 2.0])
    want_prompts = np.array(["p0", "p1", "p0", "p1", "p1", "p0", "p1", "p0"])
    from ddpo_b200.utils.stat_tracking import PerPromptStatTracker
    full = PerPromptStatTracker(32, 2).update(want_prompts, want_rewards)
    z = (want_rewards - want_rewards.mean()) / want_rewards.std()
    for i in range(WORLD):
        np.testing.assert_array_equal(r[i]["rewards"], want_rewards)
        assert list(r[i]["prompts"]) == list(want_prompts)
        np.testing.assert_allclose(r[i]["mine"], np.asarray(full).reshape(WORLD, -1)[i])
        np.testing.assert_allclose(r[i]["glob"], z.reshape(WORLD, -1)[i])
        assert list(r[i]["decoded"]) == [f"a {w} riding a bike" for w in ("zebra", "llama", "zebra", "yak", "otter", "zebra", "emu", "emu")]
    np.testing.assert_array_equal(r[0]["adv_tok"], r[1]["adv_tok"])     # same per-prompt grouping on every rank


# ------------------------------------------------------------------ the whole epoch driver on two ranks ----
@pytest.mark.timeout(600)
def test_two_rank_ddpo_driver_dry_run_keeps_replicas_in_sync(tmp_path):
    """world_size-2 gloo run of pipeline/policy_gradient.main on the CPU ops emulator (tests/_driver_worker.py): every
    rank samples its own trajectories (seed + rank, reference utils/parser.py:177), rewards are all-gathered so both
    ranks normalise over the pod's 4 samples, the gradient is all-reduced once per optimizer update, and the two model
    replicas end bit-identical"""
    import subprocess
    import sys
    port = str(_free_port())
    worker = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_driver_worker.py")
    procs = [subprocess.Popen([sys.executable, worker, str(r), "2", port, str(tmp_path)], stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT) for r in range(2)]
    logs = [p.communicate(timeout=560)[0].decode(errors="replace") for p in procs]
    assert all(p.returncode == 0 for p in procs), "\n".join(l[-1500:] for l in logs)
    r = [np.load(tmp_path / f"driver_rank{i}.npz", allow_pickle=True) for i in range(2)]
    assert int(r[0]["samples"]) == 4 and int(r[0]["step"]) == int(r[1]["step"]) == 2     # 1 minibatch / epoch / rank
    np.testing.assert_array_equal(r[0]["params"], r[1]["params"])                        # replicas in sync
    assert not np.array_equal(r[0]["rewards"], r[1]["rewards"])                          # different samples per rank
    np.testing.assert_allclose(r[0]["mean_reward"], r[1]["mean_reward"])                 # same pod statistics
    for i in range(2):
        assert float(r[i]["kl0"][0]) == 0.0                                              # first pass: ratio == 1
    # info is pmean-ed: both ranks report the same numbers
    np.testing.assert_allclose(r[0]["loss0"], r[1]["loss0"], rtol=1e-6)
    np.testing.assert_allclose(r[0]["kl1"], r[1]["kl1"], rtol=1e-6)
    assert float(r[0]["kl1"][0]) == 0.0       # epoch 1 re-samples with the updated policy: its first pass is on-policy again
    from ddpo_b200 import unet_spec
    init = unet_spec.init_flat_params(unet_spec.TINY, 0).numpy()
    assert np.abs(r[0]["params"] - init).max() > 1e-5                                    # the synchronised updates moved it
