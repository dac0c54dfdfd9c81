"""CPU, world_size 2, gloo: the host-side logic of the N>1 path - frame sharding by rank
and the single all-reduce of the flat gradient bucket.  The kernels themselves need a GPU; here only
the bucket construction / collective / ordering is exercised."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import sys
        sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
        import gd_mae_b200  # noqa: F401
        from gd_mae_b200 import config, _lib
        from gd_mae_b200.trainer import MAETrainer, optimised_parameter_names
        import bench
        from oracle import gdmae_oracle as O

        torch.manual_seed(0)  # identical initial weights on every rank
        cfg = config.builtin_cfg("tiny")
        model = config.build_mae_model(cfg)
        tr = MAETrainer(model, cfg.OPTIMIZATION, total_steps=10, world_size=world)
        # flat bucket: [optimised | never-optimised], parameters and grads are views of it
        names = [n for n, _ in model.named_parameters()]
        opt = optimised_parameter_names(model)
        assert tr.n_params == sum(p.numel() for p in model.parameters()) and tr.n_all >= tr.n_params
        for n, p in model.named_parameters():
            off, k = tr.slices[n]
            assert off % 64 == 0  # 256-byte aligned: kernels read parameters with 128-bit loads
            assert p.data.data_ptr() == tr.flat_params[off:off + k].data_ptr()
            assert p.grad.data_ptr() == tr.flat_grads[off:off + k].data_ptr()
            assert (off < tr.n_opt) == (n in opt)
        # rank-dependent gradients -> one all-reduce -> identical SUM on every rank
        for i, (n, p) in enumerate(model.named_parameters()):
            p.grad.fill_(float(rank + 1) * (1 + (i % 7)))
        tr.reduce_gradients()
        for i, (n, p) in enumerate(model.named_parameters()):
            assert torch.all(p.grad == float(sum(r + 1 for r in range(world))) * (1 + (i % 7)))
        # the CUDA-only part refuses to run on CPU (no fallback)
        try:
            tr.optimizer_step()
            raised = False
        except _lib.GdmaeError:
            raised = True
        assert raised
        # frame sharding: rank r trains on frames 1000 r + i - disjoint across ranks
        seeds = [1000 * rank + i for i in range(bench.B_PER_GPU)]
        all_seeds = [None] * world
        dist.all_gather_object(all_seeds, seeds)
        flat = [s for ss in all_seeds for s in ss]
        assert len(set(flat)) == len(flat)
        b = bench.make_batches(1, rank, O.make_cfg("tiny"), O)[0]
        assert b.shape[1] == 6 and set(b[:, 0].astype(int).tolist()) == set(range(bench.B_PER_GPU))
        q.put((rank, "ok", names[:2]))
    except Exception as e:  # pragma: no cover
        q.put((rank, repr(e), None))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_flat_bucket_allreduce_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(r[1] == "ok" for r in res), res
