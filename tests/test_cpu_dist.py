"""not gpu: the N>1 host logic on CPU with the gloo backend, world_size 2.  Each rank evaluates the
oracle's data gradient on its shard of the clip batch; the packed buffer is sum-all-reduced exactly as
FlickerAttack.step does; rank 0 checks it against the single-process full-batch gradient."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, loss_kind, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    torch.set_num_threads(2)
    from flickering_adversarial_video_b200 import dist as fdist, synthetic
    from oracle import oracle_i3d as O
    r, _, w = fdist.init_from_env("gloo")
    assert (r, w) == (rank, world)
    T, GB = 16, 2
    weights = synthetic.i3d_weights(0)
    model = O.OracleI3D(weights)
    clips = synthetic.clips_u8(GB, T, seed=1001)
    labels_all = torch.tensor([156, 156])
    delta = synthetic.delta_uniform(T, seed=7, lo=-0.05, hi=0.05)
    lo, hi = fdist.shard_range(GB, rank, world)
    x = O.normalize_u8(clips[lo:hi])
    d = delta.clone().requires_grad_(True)
    logits = model.forward(O.apply_flicker(x, d))
    if loss_kind == "improve":
        loss, _, _ = O.improve_adversarial_loss(logits, labels_all[lo:hi])          # SUM over samples
    else:
        ce, _, _ = O.ce_adversarial_loss(logits, labels_all[lo:hi])                  # local mean ...
        loss = ce * (hi - lo) / fdist.ce_grad_divisor(hi - lo, world)              # ... -> share of the global mean
    (g,) = torch.autograd.grad(loss, d)
    comm = torch.cat([g.reshape(-1), loss.detach().reshape(1)])
    fdist.allreduce_sum_(comm)
    if rank == 0:
        ret["grad"] = comm[:-1].reshape(T, 3).clone()
        ret["loss"] = float(comm[-1])
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("loss_kind", ["improve", "ce"])
def test_sharded_gradient_equals_full_batch(loss_kind):
    from flickering_adversarial_video_b200 import synthetic
    from oracle import oracle_i3d as O
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, _free_port(), loss_kind, ret), nprocs=2, join=True)
    T = 16
    model = O.OracleI3D(synthetic.i3d_weights(0))
    x = O.normalize_u8(synthetic.clips_u8(2, T, seed=1001))
    labels = torch.tensor([156, 156])
    d = synthetic.delta_uniform(T, seed=7, lo=-0.05, hi=0.05).requires_grad_(True)
    logits = model.forward(O.apply_flicker(x, d))
    loss = O.improve_adversarial_loss(logits, labels)[0] if loss_kind == "improve" else O.ce_adversarial_loss(logits, labels)[0]
    (g,) = torch.autograd.grad(loss, d)
    assert abs(ret["loss"] - float(loss)) < 1e-5 * max(1.0, abs(float(loss)))
    rel = float((ret["grad"] - g.reshape(T, 3)).norm() / g.norm())
    assert rel < 1e-2, rel      # fp32 summation order (and a few ReLU masks) differ between the sharded and the full-batch run


def _sparse_worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    torch.set_num_threads(2)
    from flickering_adversarial_video_b200 import dist as fdist, synthetic
    from oracle import oracle_resnet as R
    fdist.init_from_env("gloo")
    T, GB = 4, 2
    model = synthetic.resnet_model("r3d_18", seed=0)
    clips = synthetic.clips_u8(GB, T, 112, 112, seed=1200)
    labels = torch.tensor([3, 7])
    delta = (torch.rand((T, 112, 112, 3), generator=torch.Generator().manual_seed(5)) - 0.5) * 0.1
    lo, hi = fdist.shard_range(GB, rank, world)
    ref = R.sparse_attack_step(model, clips[lo:hi], labels[lo:hi], delta, lambda_=1.0, max_norm=0.2)
    g = ref["grad_data"].clone()                       # data term only; SparseAttack adds the L1,2 term after the exchange
    sc = torch.tensor([ref["adv_loss"]])
    fdist.allreduce_sum_(g)
    fdist.allreduce_sum_(sc)
    if rank == 0:
        ret["grad"], ret["adv_loss"] = g, float(sc)
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_per_pixel_gradient_equals_full_batch():
    """the exchange SparseAttack.step performs when world > 1 (per-pixel gradient + adversarial loss, both plain sums)"""
    from flickering_adversarial_video_b200 import synthetic
    from oracle import oracle_resnet as R
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_sparse_worker, args=(2, _free_port(), ret), nprocs=2, join=True)
    T = 4
    model = synthetic.resnet_model("r3d_18", seed=0)
    clips = synthetic.clips_u8(2, T, 112, 112, seed=1200)
    delta = (torch.rand((T, 112, 112, 3), generator=torch.Generator().manual_seed(5)) - 0.5) * 0.1
    full = R.sparse_attack_step(model, clips, torch.tensor([3, 7]), delta, lambda_=1.0, max_norm=0.2)
    assert abs(ret["adv_loss"] - full["adv_loss"]) < 1e-5 * max(1.0, abs(full["adv_loss"]))
    rel = float((ret["grad"] - full["grad_data"]).norm() / full["grad_data"].norm())
    assert rel < 1e-4, rel


def _counts_worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    from flickering_adversarial_video_b200 import dist as fdist
    fdist.init_from_env("gloo")
    got = fdist.sum_counts((3 + rank, 10 * (rank + 1)))        # rank 0: 3 of 10 fooled, rank 1: 4 of 20
    # result files / summaries of a sharded run are written by rank 0 only; per-rank replicas (world 1) all write
    assert fdist.is_writer(world=2) == (rank == 0) and fdist.is_writer(world=1) and fdist.is_writer() == (rank == 0)
    ret[rank] = got
    dist.barrier()
    dist.destroy_process_group()


def test_validation_counts_are_summed_over_ranks():
    """kinetics_i3d.evaluate / VideoLearnerAdversarial.fit take the fooling ratio over all ranks' shards"""
    from flickering_adversarial_video_b200 import dist as fdist
    assert fdist.sum_counts((3, 10)) == [3.0, 10.0]            # no process group: unchanged
    assert fdist.is_writer() and fdist.is_writer(world=1)
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_counts_worker, args=(2, _free_port(), ret), nprocs=2, join=True)
    assert ret[0] == ret[1] == [7.0, 30.0]


def _replica_worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    from flickering_adversarial_video_b200 import dist as fdist
    fdist.init_from_env("gloo")
    same = torch.arange(48, dtype=torch.float32).reshape(16, 3) * 0.01 - 0.2
    ok = fdist.replicas_equal(same)
    drift = same.clone()
    if rank == 1:
        drift[5, 2] += 1e-7                                   # one ulp-scale difference on one rank
    bad = fdist.replicas_equal(drift)
    ret[rank] = (ok, bad)
    dist.barrier()
    dist.destroy_process_group()


def test_replica_divergence_is_detected():
    from flickering_adversarial_video_b200 import dist as fdist
    assert fdist.replicas_equal(torch.zeros(3)) is True       # no process group
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_replica_worker, args=(2, _free_port(), ret), nprocs=2, join=True)
    assert ret[0] == ret[1] == (True, False)


def _attack_worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    torch.set_num_threads(1)
    from flickering_adversarial_video_b200 import attack, dist as fdist
    from test_cpu_attack_host_logic import K, T, StandInEngine, _data
    from test_cpu_host_flows_standin import SparseStandIn
    fdist.init_from_env("gloo")
    clips, delta, labels = _data()                                     # 2 clips: one per rank
    lo, hi = fdist.shard_range(2, rank, world)
    attack.FlickerEngine = StandInEngine
    atk = attack.FlickerAttack({}, 1, T, {"LAMBDA": 1.0, "BETA_1": 0.5}, num_classes=K, arch="r3d_18", delta_clip=0.1)
    assert atk.world == 2 and atk.global_batch == 2
    atk.delta.copy_(delta)
    for _ in range(3):
        atk.step(clips[lo:hi], labels[lo:hi])
    atk.check_replicas()
    solo = attack.FlickerAttack({}, 1, T, {}, num_classes=K, arch="r3d_18", sharded=False)
    assert solo.world == 1
    attack.FlickerEngine = SparseStandIn
    torch.manual_seed(100 + rank)                                      # ranks would draw different initial perturbations ...
    sp = attack.SparseAttack({}, 1, T, {"LAMBDA": 1.0}, num_classes=K, arch="r3d_18")
    assert sp.world == 2
    sp.check_replicas()                                                # ... the constructor broadcasts rank 0's
    for _ in range(2):
        sp.step(clips[lo:hi], labels[lo:hi])
    sp.check_replicas()
    if rank == 0:
        ret["delta"], ret["adv"], ret["sparse"] = atk.delta.clone(), float(atk.scalars[0]), sp.delta.clone()
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_attack_objects_equal_the_full_batch():
    """attack.FlickerAttack / SparseAttack with world_size 2 (gloo) on the stand-in engine: the packed all-reduce inside
    step() makes two one-clip ranks walk the same trajectory as one process with both clips"""
    from flickering_adversarial_video_b200 import attack
    from test_cpu_attack_host_logic import K, T, StandInEngine, _data
    from test_cpu_host_flows_standin import SparseStandIn
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_attack_worker, args=(2, _free_port(), ret), nprocs=2, join=True)
    clips, delta, labels = _data()
    saved = attack.FlickerEngine
    try:
        attack.FlickerEngine = StandInEngine
        full = attack.FlickerAttack({}, 2, T, {"LAMBDA": 1.0, "BETA_1": 0.5}, num_classes=K, arch="r3d_18", delta_clip=0.1)
        full.delta.copy_(delta)
        for _ in range(3):
            full.step(clips, labels)
        attack.FlickerEngine = SparseStandIn
        torch.manual_seed(0)
        sp = attack.SparseAttack({}, 2, T, {"LAMBDA": 1.0}, num_classes=K, arch="r3d_18", init=torch.zeros((T, 12, 12, 3)))
    finally:
        attack.FlickerEngine = saved
    assert torch.allclose(ret["delta"], full.delta, rtol=1e-4, atol=1e-7)
    assert abs(ret["adv"] - float(full.scalars[0])) < 1e-5 * max(1.0, abs(ret["adv"]))
    assert ret["sparse"].shape == sp.delta.shape


def test_shard_range_rules():
    from flickering_adversarial_video_b200 import dist as fdist
    assert fdist.shard_range(64, 3, 8) == (24, 32)
    with pytest.raises(ValueError):
        fdist.shard_range(10, 0, 4)
    assert fdist.ce_grad_divisor(8, 8) == 64


def _lockstep_worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    from flickering_adversarial_video_b200 import dist as fdist
    fdist.init_from_env("gloo")
    # unequal shards: rank 0 can supply 5 batches, rank 1 only 3 (record files of different length)
    n_local = 5 if rank == 0 else 3
    for mode in ("count", "flag"):
        steps = 0
        src = ((rank, i) for i in range(n_local))
        for _ in fdist.lockstep(src, world, num_batches=n_local if mode == "count" else None):
            t = torch.ones(1)
            fdist.allreduce_sum_(t)            # the step's collective: must pair up on every rank
            assert float(t) == world
            steps += 1
        c = fdist.sum_counts((steps,))         # the collective that FOLLOWS the loop (validation counts)
        ret[(rank, mode)] = (steps, c[0])
    # the split permutation of rank 0 reaches every rank; shards are disjoint and of equal length
    perm = torch.randperm(11, generator=torch.Generator().manual_seed(100 + rank)).tolist()
    same = fdist.broadcast_ints(perm)
    ret[(rank, "perm")] = same
    ret[(rank, "shard")] = fdist.shard_indices(same, rank, world)
    dist.barrier()
    dist.destroy_process_group()


def test_unequal_shards_stay_in_lockstep():
    """ranks whose record files hold different numbers of batches must run the same number of steps (each
    step is a collective) before they move on to the validation pass's count exchange."""
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_lockstep_worker, args=(2, _free_port(), ret), nprocs=2, join=True)
    for mode in ("count", "flag"):
        assert ret[(0, mode)] == (3, 6.0) and ret[(1, mode)] == (3, 6.0), dict(ret)
    assert ret[(0, "perm")] == ret[(1, "perm")]
    a, b = ret[(0, "shard")], ret[(1, "shard")]
    assert len(a) == len(b) == 5 and not set(a) & set(b)
