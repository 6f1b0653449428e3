"""The named synthetic workloads of BASELINE.json / SURVEY.md 8(d) as line descriptors: the reference's own
example models and the configurations C3-C5.  Used by bench.py, tools/ and the tests (tests/lines.py)."""
from __future__ import annotations

import numpy as np

from . import abi

SEED = 0x4D4A50646573  # "MJPdes"


def example_line(warm_up=2000.0, run_to=20000.0, **kw) -> abi.HostLine:
    """resources/example.clj:2-19 (5 machines, 4 buffers, 2 job types)."""
    return abi.HostLine(lam=[0.1] * 5, mu=[0.9] * 5, W=[1.2, 1.0, 1.1, 1.05, 1.2], discipline=[0] * 5,
                        N=[3, 5, 1, 1], portion=[0.8, 0.2],
                        w=[[1.0, 1.0, 1.0, 1.0, 1.0], [1.0, 2.0, 1.5, 1.0, 1.0]],
                        warm_up=warm_up, run_to=run_to, **kw)


def example2_line(warm_up=2000.0, run_to=20000.0, **kw) -> abi.HostLine:
    """resources/example2.clj (single job type)."""
    return abi.HostLine(lam=[0.1] * 5, mu=[0.9] * 5, W=[1.2, 1.0, 1.1, 1.05, 1.2], discipline=[0] * 5,
                        N=[3, 5, 1, 1], portion=[1.0], w=[[1.0] * 5], warm_up=warm_up, run_to=run_to, **kw)


def submodel1_line(warm_up=2000.0, run_to=20000.0, **kw) -> abi.HostLine:
    """data/submodel-1.clj (2 machines, N=3)."""
    return abi.HostLine(lam=[0.1] * 2, mu=[0.9] * 2, W=[1.0, 1.0], discipline=[0, 0], N=[3], portion=[1.0],
                        w=[[1.0, 1.0]], warm_up=warm_up, run_to=run_to, **kw)


def sweep10_line(n_inst: int, warm_up=2000.0, run_to=20000.0, first=0, total=1_000_000) -> abi.HostLine:
    """C3: 10-machine line, 4-axis grid 10 x 10 x 100 x 100 over (N_all, N_b5, W_all, W_m5);
    instance id = point index, this call holds points [first, first+n_inst)."""
    M = 10
    idx = np.arange(first, first + n_inst, dtype=np.int64)
    i_w5 = idx % 100
    i_w = (idx // 100) % 100
    i_n5 = (idx // 10_000) % 10
    i_n = (idx // 100_000) % 10
    grid = np.linspace(0.8, 1.2, 100)
    W = np.tile(grid[i_w], (M, 1))
    W[4] = grid[i_w5]
    N = np.tile((1 + i_n).astype(np.int32), (M - 1, 1))
    N[4] = (1 + i_n5).astype(np.int32)
    w = np.ones((1, M))
    w[0, M - 1] = 1.01
    return abi.HostLine(lam=[0.1] * M, mu=[0.9] * M, W=[1.0] * M, discipline=[0] * M, N=[10] * (M - 1),
                        portion=[1.0], w=w, warm_up=warm_up, run_to=run_to, W_inst=W, N_inst=N)


def mixed20_line(warm_up=2000.0, run_to=20000.0, log=True, up_down=True, max_lines=abi.UINT64_MAX) -> abi.HostLine:
    """C4: 20 machines, BBS on the even-numbered names (m2, m4, ... = odd indices), 8 job types, SCADA on."""
    M, K = 20, 8
    i = np.arange(M)
    return abi.HostLine(lam=[0.1] * M, mu=[0.9] * M, W=1.0 + 0.05 * ((7 * i) % 5), discipline=(i % 2 == 1),
                        N=1 + (np.arange(M - 1) % 5), portion=[0.125] * K,
                        w=[[1.0 + 0.25 * ((k + j) % 4) for j in range(M)] for k in range(K)],
                        warm_up=warm_up, run_to=run_to, log=log, up_down=up_down, max_lines=max_lines)


def long50_line(warm_up=1e5, run_to=1e6) -> abi.HostLine:
    """C5: 50 machines, 16 job types."""
    M, K = 50, 16
    i = np.arange(M)
    return abi.HostLine(lam=[0.1] * M, mu=[0.9] * M, W=1.0 + 0.02 * ((11 * i) % 11), discipline=[0] * M,
                        N=1 + (np.arange(M - 1) % 4), portion=[1.0 / K] * K,
                        w=[[1.0 + 0.125 * ((3 * k + 5 * j) % 8) for j in range(M)] for k in range(K)],
                        warm_up=warm_up, run_to=run_to)
