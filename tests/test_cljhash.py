"""Clojure map iteration order (mjpdes_b200/cljhash.py, oracle/mjp_oracle.cpp, csrc/mjp_plan.h): the three
restatements agree, reproduce the published known value, and drive the places where the reference walks a map of
more than 8 entries: clean-log-buf's (group-by :m) (util/log.clj:44), new-job's walk over :jobmix (core.clj:332)
and calc-bneck's (vals (:blocked res)) (core.clj:641-642)."""
import numpy as np

from hostsim import binding as hostsim
from lines import SEED, mixed20_line, sweep10_line
from mjpdes_b200 import abi, cljhash, model
from parity import assert_same


def test_known_keyword_hash():
    # (hash :a) in Clojure >= 1.6 (Murmur3-based hasheq) -- the value quoted in Clojure's own hashing notes
    assert cljhash.to_signed(cljhash.keyword_hash(":a")) == -2123407586


def test_python_and_oracle_restatements_agree(oracle):
    lib = oracle.lib()
    for name in ["a", "m1", "m2", "m10", "m33", "m64", "b7", "jobType1", "jobType16", "x", "xy", "xyz"]:
        h = cljhash.keyword_hash(":" + name)
        assert lib.mjpo_clj_keyword_hash(name.encode()) == h
        key = 0
        for c in cljhash.trie_key(h):
            key = (key << 5) | c
        assert lib.mjpo_clj_trie_key(h) == key


def test_map_order_rules():
    small = [f":m{i}" for i in range(1, 9)]
    assert cljhash.map_order(small) == small                       # array-map: insertion order up to 8 entries
    big = [f":m{i}" for i in range(1, 11)]
    order = cljhash.map_order(big)
    assert sorted(order) == sorted(big) and order != big            # 9th key promotes to a hash map
    assert order == cljhash.map_order(big[::-1])                    # ... whose order ignores insertion order
    rank = cljhash.hash_rank(big)
    assert [big[i] for i in np.argsort(rank)] == order


def _logged(line, n, cap):
    line.log, line.up_down, line.max_lines = True, True, abi.UINT64_MAX
    return cap * n


def test_wide_ticks_are_regrouped_in_hash_order(oracle):
    """A lattice line (identical machines) has ticks in which more than 8 machines log; the kernel source (host
    shim) and the oracle regroup them identically -- statistics and SCADA records -- both for the default names
    :m1.. and for an explicit group_rank, and the regrouping is NOT the first-appearance one."""
    n = 6
    line = sweep10_line(n, warm_up=100, run_to=1500, first=5050)
    cap = _logged(line, n, 60000)
    want = oracle.run(line, n, seed=SEED, log_capacity=cap)
    assert want.extra["wide_ticks"].sum() > 100 and want.extra["wide_ticks_reordered"].sum() > 100
    got = hostsim.run(line, n, seed=SEED, log_capacity=cap)
    assert_same(got, want.stats)
    key = lambda r: np.argsort((r["instance"].astype(np.uint64) << np.uint64(32)) | r["line"].astype(np.uint64))
    a, b = got.log[key(got.log)], want.log[key(want.log)]
    for f in ("instance", "line", "act", "machine", "n", "job", "clk"):
        assert np.array_equal(a[f], b[f]), f
    # other machine names -> another order -> another record stream, still identical on both sides
    line2 = mixed20_line(warm_up=100, run_to=1500)
    line2.group_rank = np.arange(20, dtype=np.int32)[::-1].copy()
    w2 = oracle.run(line2, n, seed=SEED, log_capacity=60000 * n)
    g2 = hostsim.run(line2, n, seed=SEED, log_capacity=60000 * n)
    assert_same(g2, w2.stats)
    a, b = g2.log[key(g2.log)], w2.log[key(w2.log)]
    assert np.array_equal(a["machine"], b["machine"]) and np.array_equal(a["act"], b["act"])
    w_default = oracle.run(mixed20_line(warm_up=100, run_to=1500), n, seed=SEED, log_capacity=60000 * n)
    assert not np.array_equal(w_default.log[key(w_default.log)]["machine"], b["machine"])


def _model(M, K):
    names = [f":m{i}" for i in range(1, M + 1)]
    topo, line = [], {}
    for i, m in enumerate(names):
        topo.append(m)
        line[m] = {":lambda": 0.1, ":mu": 0.9, ":W": 1.0}
        if i < M - 1:
            topo.append(f":b{i + 1}")
            line[f":b{i + 1}"] = {":N": 2}
    jm = {f":jobType{k + 1}": {":portion": 1.0 / K, ":w": {m: 1.0 + 0.1 * k for m in names}} for k in range(K)}
    return {":line": line, ":topology": topo, ":entry-point": ":m1", ":jobmix": jm,
            ":params": {":warm-up-time": 10, ":run-to-time": 100}}


def test_jobmix_is_walked_in_map_order():
    cl = model.preprocess_model(_model(3, 8), n_inst=1)
    assert cl.jobtypes == [f":jobType{k}" for k in range(1, 9)]              # <= 8: as written
    cl = model.preprocess_model(_model(3, 12), n_inst=1)
    assert cl.jobtypes == cljhash.map_order([f":jobType{k}" for k in range(1, 13)]) != [f":jobType{k}" for k in range(1, 13)]
    # row k of w belongs to jobtypes[k]
    for k, jt in enumerate(cl.jobtypes):
        assert cl.host.w[k, 0] == 1.0 + 0.1 * (int(jt[len(":jobType"):]) - 1)
    assert list(cl.host.group_rank) == cljhash.hash_rank([":m1", ":m2", ":m3"])


def test_calc_bneck_beyond_eight_machines_follows_the_reference_quirk():
    """M = 10: (vals (:blocked res)) is in hash order, so the reference's bl(i)/st(i) are not machine i's figures
    (SURVEY Q10).  :bottleneck-machine reproduces that; :bottleneck-machine-by-position is the intended rule."""
    machines = [f":m{i}" for i in range(1, 11)]
    rng = np.random.default_rng(5)
    res = {":blocked": {m: float(x) for m, x in zip(machines, rng.uniform(0, .3, 10))},
           ":starved": {m: float(x) for m, x in zip(machines, rng.uniform(0, .3, 10))}}
    out = model.calc_bneck(res, machines)
    order = cljhash.map_order(machines)
    bl = [res[":blocked"][m] for m in order]
    st = [res[":starved"][m] for m in order]
    cand = [i for i in range(9) if bl[i] < st[i + 1]]
    if len(cand) == 1:
        expect = machines[cand[0]]
    else:
        sev = [abs(bl[0] - st[1])] + [abs(bl[i - 1] - st[i]) + abs(bl[i] - st[i + 1]) for i in range(1, 9)] + [abs(bl[8] - st[9])]
        top = max(sev)
        expect = next(m for m in order if sev[machines.index(m)] == top)
    assert out[":bottleneck-machine"] == expect
    assert out[":bottleneck-machine-by-position"] == model.calc_bneck(res, machines, faithful=False)[":bottleneck-machine"]
    # up to 8 machines both readings coincide and no extra key appears
    m8 = machines[:8]
    r8 = {k: {m: v[m] for m in m8} for k, v in res.items()}
    assert ":bottleneck-machine-by-position" not in model.calc_bneck(r8, m8)
