"""Training-path plumbing at N>1 on CPU (BASELINE config 4): two gloo ranks train a small torch model on different data
through rdmnet_b200.ddp.BucketedGradAllReduce; the averaged gradients must equal the gradients of the concatenated batch
(what torch DistributedDataParallel gives the reference, geotransformer/engine/base_trainer.py:181-191), parameters that
receive no gradient on one rank must not dead-lock the collectives, and the logged scalars travel in ONE all-reduce."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from rdmnet_b200 import ddp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _model():
    torch.manual_seed(5)
    return torch.nn.Sequential(torch.nn.Linear(12, 32), torch.nn.ReLU(), torch.nn.Linear(32, 32), torch.nn.ReLU(), torch.nn.Linear(32, 4))


def _data(rank):
    g = torch.Generator().manual_seed(100 + rank)
    return torch.randn(16, 12, generator=g), torch.randn(16, 4, generator=g)


def _worker(rank, world, port, comm_dtype, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        model = _model()
        unused = torch.nn.Parameter(torch.ones(7))  # a parameter no rank ever touches (the reference logs these: epoch_based_trainer.py:105-107)
        model.register_parameter("unused", unused)
        red = ddp.BucketedGradAllReduce(model, bucket_bytes=512, comm_dtype=comm_dtype)
        assert len(red.buckets) >= 2
        grads = []
        for step in range(2):  # two steps: the reducer must re-arm
            x, y = _data(rank + 10 * step)
            if step == 0:
                red.zero_grad()
            else:
                model.zero_grad(set_to_none=True)  # the habit the reducer must survive: gradient tensors are re-created by autograd
            out = model(x)
            if rank == 1 and step == 1:
                loss = ((out[:, :2] - y[:, :2]) ** 2).mean() * 0.5  # rank-dependent graph: still no dead-lock
            else:
                loss = ((out - y) ** 2).mean()
            loss.backward()
            red.finish()
            grads.append({n: p.grad.clone() for n, p in model.named_parameters() if p.grad is not None})
        scal = ddp.all_reduce_scalars({"loss": loss.detach(), "rank": float(rank), "const": 3.0})
        q.put((rank, [{k: v.numpy() for k, v in g.items()} for g in grads], {k: float(v) for k, v in scal.items()}, red.bytes_per_step))
    finally:
        dist.destroy_process_group()


def _run(comm_dtype):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, comm_dtype, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = sorted((q.get(timeout=120) for _ in range(world)), key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    return got


def _expected(step):
    model = _model()
    total = 0
    for rank in range(2):
        x, y = _data(rank + 10 * step)
        out = model(x)
        if rank == 1 and step == 1:
            total = total + ((out[:, :2] - y[:, :2]) ** 2).mean() * 0.5
        else:
            total = total + ((out - y) ** 2).mean()
    (total / 2).backward()
    return {n: p.grad for n, p in model.named_parameters()}


def test_bucketed_allreduce_fp32_equals_mean_of_rank_gradients():
    got = _run(torch.float32)
    for step in range(2):
        exp = _expected(step)
        for rank in range(2):
            g = got[rank][1][step]
            for n, e in exp.items():
                assert torch.allclose(torch.from_numpy(g[n]), e, rtol=1e-5, atol=1e-7), (step, rank, n)
            assert "unused" in g and not g["unused"].any()  # never touched: stays zero on every rank
    assert got[0][2] == got[1][2]
    assert abs(got[0][2]["rank"] - 0.5) < 1e-6 and abs(got[0][2]["const"] - 3.0) < 1e-6


def test_bucketed_allreduce_bf16_wire_format():
    got = _run(torch.bfloat16)
    exp = _expected(0)
    for rank in range(2):
        g = got[rank][1][0]
        for n, e in exp.items():  # bf16 on the wire: 8 mantissa bits per addend
            assert torch.allclose(torch.from_numpy(g[n]), e, rtol=2e-2, atol=2e-3 * float(e.abs().max())), (rank, n)
        for n in exp:  # both ranks hold the SAME reduced values (the optimizer states must not diverge)
            assert (got[0][1][0][n] == got[1][1][0][n]).all()
    n_par = sum(p.numel() for p in _model().parameters()) + 7
    assert got[0][3] == 2 * n_par
