"""World-size-2 test of the sharded round loop (basq_b200.sharded.recombine_sharded) on CPU with
gloo.  The CUDA session is replaced by an oracle-backed engine with the same stage interface, so
this exercises the host protocol: global set assignment by offset, the single all-reduce per round,
analytic survivor counts on every rank, termination, and the final gather."""
import math
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from basq_b200 import sharded
from oracle import gp_kernels as ogp
from oracle import rchq as orchq


class OracleEngine:
    """count / partial / car / apply / result with torch CPU ops (tests only)."""

    def __init__(self, X_loc, Z, U, kernel, N_glob, idx_base, mu_loc=None):
        self.X, self.Z, self.U, self.kernel = X_loc.double(), Z.double(), U.double(), kernel
        self.mu = torch.full((len(X_loc),), 1.0 / N_glob, dtype=torch.float64) if mu_loc is None else mu_loc.double()
        self.idx = torch.arange(len(X_loc)) + idx_base
        live = self.mu != 0
        self.X, self.mu, self.idx = self.X[live], self.mu[live], self.idx[live]
        self.n = U.shape[0] + 1
        self.S = 2 * self.n
        self.device = "cpu"

    def count(self):
        return len(self.mu)

    def cell_factor(self, R, R_loc_max):
        F = 1
        while F < 4 and R >= 4 * (2 * F) * self.S:
            F *= 2
        return F

    def _cells(self, off, F):
        return (off + torch.arange(len(self.mu))) % (F * self.S)

    def pass_begin(self, R, off, F):
        self.F = F
        self.cellA = torch.zeros(self.n, F * self.S, dtype=torch.float64)
        if len(self.mu) == 0:
            return
        cells = self._cells(off, F)
        feats = (self.U @ self.kernel(self.Z, self.X)) * self.mu.unsqueeze(0)       # [q, R_loc]
        self.cellA[0].index_add_(0, cells, self.mu)
        self.cellA[1:].index_add_(1, cells, feats)

    def level(self, lvl, node, ppos, fpar, A):
        """Columns of the level's nodes, every one folded directly from the cell columns (the
        library derives the high halves as parent - low half; the two must agree)."""
        stride, cnt = self.S << lvl, self.F >> lvl
        ids = [int(u) for u in node] + ([int(u) + (self.S << (lvl - 1)) for u in node] if lvl > 0 else [])
        fs = [float(f) for f in fpar] * (2 if lvl > 0 else 1)
        A.zero_()
        for i, (u, f) in enumerate(zip(ids, fs)):
            A[:, i] = f * sum(self.cellA[:, u + k * stride] for k in range(cnt))

    def car(self, A, C, omega):
        mass = A[0, :C]
        bary = (A[1:, :C] / mass.unsqueeze(0)).T
        w, keep = orchq.caratheodory(bary, mass.clone())
        omega.zero_()
        omega[keep] = w / mass[keep]

    def apply(self, R, off, F, factor):
        scale = factor[self._cells(off, F)]
        live = scale > 0
        self.X, self.mu, self.idx = self.X[live], (self.mu * scale)[live], self.idx[live]
        return len(self.mu)

    def result(self):
        return self.idx.clone(), self.mu.clone()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, N, weighted, out_path):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    os.environ.setdefault("GLOO_SOCKET_IFNAME", "lo")
    # file-based rendezvous: no race for a TCP port between bind-and-release and the store's own bind
    dist.init_process_group("gloo", init_method=f"file://{out_path}.rdzv", rank=rank, world_size=world)
    try:
        torch.set_num_threads(2)
        g = torch.Generator().manual_seed(123)
        d, M, n = 3, 30, 7
        X = math.sqrt(2.0) * torch.randn(N, d, generator=g, dtype=torch.float64)
        Z = math.sqrt(2.0) * torch.randn(M, d, generator=g, dtype=torch.float64)
        U = torch.linalg.qr(torch.randn(M, n - 1, generator=g, dtype=torch.float64)).Q.T.contiguous()
        mu = None
        if weighted:
            mu = torch.rand(N, generator=g, dtype=torch.float64)
            mu[torch.rand(N, generator=g) < 0.3] = 0.0
            mu = mu / mu.sum()
        kern = lambda a, b: ogp.base_kernel(a, b, "rbf", 1.4, 1.0)
        lo, hi = sharded.shard_bounds(N, world, rank)
        eng = OracleEngine(X[lo:hi], Z, U, kern, N, lo, None if mu is None else mu[lo:hi])
        idx, w = sharded.recombine_sharded(eng, eng.n, eng.S)
        idx, w = sharded.gather_result(idx, w, eng.n)
        assert len(idx) <= n and bool((w > 0).all()) and bool((idx[1:] > idx[:-1]).all())
        full_mu = torch.full((N,), 1.0 / N, dtype=torch.float64) if mu is None else mu
        Phi = orchq.features(X, U, Z, kern)
        res = orchq.moment_residual(Phi, full_mu, idx, w)
        # fp64 oracle engine; the Caratheodory pivot order (and with it the last digits) depends on the
        # threading of the host BLAS, observed 3e-12 .. 3e-11
        assert res < 2e-10, res
        assert abs(float(w.sum()) - 1.0) < 1e-12
        # every rank holds the same gathered rule
        chk = torch.stack([idx.double().sum(), w.sum()])
        ref = chk.clone()
        dist.broadcast(ref, 0)
        assert torch.equal(chk, ref)
        if rank == 0:
            torch.save({"idx": idx, "w": w, "res": res}, out_path)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("N,weighted", [(2000, False), (1537, True), (9, False), (20, False)])
def test_two_rank_round_loop(tmp_path, N, weighted):
    out = str(tmp_path / "rule.pt")
    for attempt in range(3):     # the rendezvous port is picked by bind-and-release: retry if someone else grabbed it
        try:
            mp.spawn(_worker, args=(2, _free_port(), N, weighted, out), nprocs=2, join=True)
            break
        except Exception as e:  # noqa: BLE001
            if attempt == 2 or "AssertionError" in str(e):
                raise
    r = torch.load(out)
    assert len(r["idx"]) >= 1


def test_single_process_matches_two_rank_invariants():
    """world size 1 (no process group): same engine, same loop."""
    g = torch.Generator().manual_seed(5)
    N, d, M, n = 800, 3, 30, 7
    X = torch.randn(N, d, generator=g, dtype=torch.float64)
    Z = torch.randn(M, d, generator=g, dtype=torch.float64)
    U = torch.linalg.qr(torch.randn(M, n - 1, generator=g, dtype=torch.float64)).Q.T.contiguous()
    kern = lambda a, b: ogp.base_kernel(a, b, "rbf", 1.4, 1.0)
    eng = OracleEngine(X, Z, U, kern, N, 0)
    idx, w = sharded.recombine_sharded(eng, eng.n, eng.S)
    Phi = orchq.features(X, U, Z, kern)
    assert orchq.moment_residual(Phi, torch.full((N,), 1.0 / N, dtype=torch.float64), idx, w) < 1e-11
