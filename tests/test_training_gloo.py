"""Data-parallel training plumbing on CPU: two gloo ranks with half the batch each must produce
the gradients, the clipped global norm and the parameter update of ONE process with the whole batch (losses that are
batch means; train_fastsvc.py:199-235: clip after the reduce, then step).  The toy generator / critic are plain
torch.nn -- the plumbing under test (GradBucket, GanTrainer) does not care which modules it is given."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
from torch import nn

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "tools"))

from svcc23_fastsvc_b200.training import GanTrainer, GradBucket  # noqa: E402


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


class ToyG(nn.Module):
    def __init__(self):
        super().__init__()
        self.a = nn.Conv1d(4, 8, 3, padding=1)
        self.b = nn.Conv1d(8, 1, 3, padding=1)

    def forward(self, x, s):
        return self.b(torch.tanh(self.a(x) + s))


class ToyD(nn.Module):
    def __init__(self):
        super().__init__()
        self.c = nn.Conv1d(1, 4, 5, stride=2)
        self.o = nn.Conv1d(4, 1, 3)

    def forward(self, x):
        return [self.o(torch.relu(self.c(x)))]


def _l1_pair(y_, y):           # batch-mean stand-in for (sc_loss, mag_loss)
    return (y_ - y).abs().mean(), ((y_ - y) ** 2).mean()


def _gen_adv(outs):
    return sum(((o - 1) ** 2).mean() for o in outs) / len(outs)


def _dis_adv(outs_hat, outs):
    return (sum(((o - 1) ** 2).mean() for o in outs) / len(outs),
            sum((o ** 2).mean() for o in outs_hat) / len(outs_hat))


def _data(n=8, T=64):
    g = torch.Generator().manual_seed(5)
    return torch.randn(n, 4, T, generator=g), torch.randn(n, 1, T, generator=g), torch.randn(n, 1, T, generator=g)


def _models():
    torch.manual_seed(11)
    return ToyG(), ToyD()


def _run_steps(G, D, x, s, y, n_steps, group_world):
    opt_g = torch.optim.SGD(G.parameters(), lr=0.05)
    opt_d = torch.optim.SGD(D.parameters(), lr=0.05)
    tr = GanTrainer(G, D, _l1_pair, _gen_adv, _dis_adv, opt_g, opt_d, lambda_adv=2.5, generator_grad_norm=0.5,
                    discriminator_grad_norm=0.25)
    logs = None
    for _ in range(n_steps):
        logs = tr.step((x, s), y, adversarial=True)
    tr.finish_discriminator_step()
    return logs


def _reference_steps(G, D, x, s, y, n_steps):
    """train_fastsvc.py:157-235 literally, one process, whole batch."""
    opt_g = torch.optim.SGD(G.parameters(), lr=0.05)
    opt_d = torch.optim.SGD(D.parameters(), lr=0.05)
    norms = []
    for _ in range(n_steps):
        y_ = G(x, s)
        a, b = _l1_pair(y_, y)
        gen_loss = a + b + 2.5 * _gen_adv(D(y_))
        opt_g.zero_grad()
        gen_loss.backward()
        norms.append(float(torch.nn.utils.clip_grad_norm_(G.parameters(), 0.5)))
        opt_g.step()
        with torch.no_grad():
            y_ = G(x, s)
        real, fake = _dis_adv(D(y_.detach()), D(y))
        opt_d.zero_grad()
        (real + fake).backward()
        torch.nn.utils.clip_grad_norm_(D.parameters(), 0.25)
        opt_d.step()
    return norms


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.set_num_threads(1)
        x, s, y = _data()
        n = x.shape[0] // world
        sl = slice(rank * n, (rank + 1) * n)
        G, D = _models()
        logs = _run_steps(G, D, x[sl], s[sl], y[sl], 3, world)
        q.put((rank, [p.detach().numpy().copy() for p in G.parameters()],          # numpy: pickled by value
               [p.detach().numpy().copy() for p in D.parameters()], float(logs["generator_grad_norm"])))
    finally:
        dist.destroy_process_group()


def test_two_rank_training_equals_single_process_whole_batch():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted((q.get(timeout=180) for _ in range(world)), key=lambda r: r[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    x, s, y = _data()
    G, D = _models()
    norms = _reference_steps(G, D, x, s, y, 3)
    for rank, gp, dp, gnorm in res:
        for a, b in zip(gp, G.parameters()):
            assert torch.allclose(torch.from_numpy(a), b, rtol=1e-5, atol=1e-6), rank
        for a, b in zip(dp, D.parameters()):
            assert torch.allclose(torch.from_numpy(a), b, rtol=1e-5, atol=1e-6), rank
        assert abs(gnorm - norms[-1]) <= 1e-5 * max(1.0, norms[-1])      # the GLOBAL norm, taken after the reduce
    for a, b in zip(res[0][1], res[1][1]):                              # replicas stay bit-identical
        assert (a == b).all()


def test_grad_bucket_single_process():
    torch.manual_seed(0)
    m = nn.Sequential(nn.Linear(5, 7), nn.Tanh(), nn.Linear(7, 3))
    ref = nn.Sequential(nn.Linear(5, 7), nn.Tanh(), nn.Linear(7, 3))
    ref.load_state_dict(m.state_dict())
    b = GradBucket(m.parameters())
    assert b.numel == sum(p.numel() for p in m.parameters()) and b.world == 1
    x = torch.randn(4, 5)
    for _ in range(2):                      # backward accumulates INTO the views; zero() resets them
        b.zero()
        m(x).pow(2).sum().backward()
    ref(x).pow(2).sum().backward()
    for p, q in zip(m.parameters(), ref.parameters()):
        assert p.grad.data_ptr() >= b.flat.data_ptr() and torch.allclose(p.grad, q.grad)
    b.all_reduce()                          # no process group: identity
    want = torch.nn.utils.clip_grad_norm_(ref.parameters(), 0.1)
    got = b.clip_(0.1)
    assert torch.allclose(got, want)
    for p, q in zip(m.parameters(), ref.parameters()):
        assert torch.allclose(p.grad, q.grad, rtol=1e-6, atol=1e-8)
    m.zero_grad(set_to_none=True)
    with pytest.raises(RuntimeError, match="no longer aliases"):
        b.all_reduce()
    with pytest.raises(ValueError):
        GradBucket([])
