"""Host-side mirror of the reference interface: EDN model files -> flat descriptor
(preprocess-model), result maps (calc-basics / calc-bneck), SCADA text (print-lines)."""
import io
import math

import numpy as np
import pytest

from lines import GOLDENS, SEED, example_line, golden_two_machine
from mjpdes_b200 import abi, edn, model, philox, scada


def as_model(d):
    """golden JSON (plain dicts) -> Model with record types, as eval of the file would give."""
    line = {k: (model.Buffer(v) if ":N" in v else model.ExpoMachine(v)) for k, v in d[":line"].items()}
    jm = {k: model.JobType(v) for k, v in d[":jobmix"].items()}
    return model.Model({**d, ":line": line, ":jobmix": jm})


# ---- EDN reader ------------------------------------------------------------------------------------

def test_edn_reads_a_model_form():
    text = """(map->Model {:line {:m1 (map->ExpoMachine {:lambda 0.1 :mu 0.9 :W 1.2 }) ; comment
                                  :b1 (map->Buffer {:N 3}) :m2 (core/map->ExpoMachine {:lambda 0 :mu 1, :W {:dist :uniform :bounds [0.8, 1.2]}})}
                 :topology [:m1 :b1 :m2] :entry-point :m1 #_(ignored form)
                 :report {:max-lines ##Inf :log? true} :params {:warm-up-time 2000 :run-to-time 20000}
                 :jobmix {:jobType1 (map->JobType {:portion 1.0 :w {:m1 1.0, :m2 1.01}})}})"""
    m = model.eval_form(edn.read_string(text))
    assert isinstance(m, model.Model) and isinstance(m[":line"][":m1"], model.ExpoMachine)
    assert m[":line"][":b1"][":N"] == 3 and m[":report"][":max-lines"] == math.inf
    assert m[":line"][":m2"][":W"][":bounds"] == [0.8, 1.2]
    assert list(m[":line"]) == [":m1", ":b1", ":m2"]
    assert edn.read_string('[1 2.5 -3 1/2 "s" nil true #{1 2}]')[:7] == [1, 2.5, -3, 0.5, "s", None, True]


def test_edn_errors():
    for bad in ("(map->Model {:a 1", "{:a}", ")"):
        with pytest.raises(edn.EdnError):
            edn.read_string(bad)


# ---- preprocess-model --------------------------------------------------------------------------------

def test_example_clj_compiles_to_the_expected_descriptor():
    cl = model.preprocess_model(as_model(GOLDENS["models"]["resources/example.clj"]))
    want = example_line()
    for f in ("lam", "mu", "W", "discipline", "N", "portion", "w"):
        assert np.array_equal(getattr(cl.host, f), getattr(want, f)), f
    assert (cl.host.warm_up, cl.host.run_to, cl.host.run_to_job) == (2000.0, 20000.0, -1)
    assert not cl.host.log and cl.host.max_lines == 0 and not cl.host.jm2dm        # core.clj:586-588,582
    assert cl.machines == [":m1", ":m2", ":m3", ":m4", ":m5"] and cl.jobtypes == [":jobType1", ":jobType2"]
    assert cl.n_inst == 1


@pytest.mark.parametrize("rel,M,K,n", [("resources/example2.clj", 5, 1, 1), ("data/submodel-1.clj", 2, 1, 20),
                                        ("data/quick.clj", 5, 2, 1), ("data/example-10.clj", 5, 1, 10)])
def test_reference_model_files(rel, M, K, n):
    cl = model.preprocess_model(as_model(GOLDENS["models"][rel]))
    assert (cl.host.M, cl.host.K, cl.n_inst) == (M, K, n)


def test_bounds_are_sampled_per_replication():
    """data/example-10.clj: W ~ U[0.8, 1.2] re-drawn for every replication (core.clj:563)."""
    m = as_model(GOLDENS["models"]["data/example-10.clj"])
    a = model.preprocess_model(m, n_inst=1000, seed=SEED)
    W = a.host.W_inst
    assert W.shape == (5, 1000) and W.min() >= 0.8 and W.max() < 1.2
    assert abs(W.mean() - 1.0) < 0.01 and abs(W.std() - 0.4 / math.sqrt(12)) < 0.01
    assert len(np.unique(W)) == W.size
    # a replication's draw depends on its instance id only: sharding the batch changes nothing
    b = model.preprocess_model(m, n_inst=500, seed=SEED, first_instance_id=500)
    assert np.array_equal(b.host.W_inst, W[:, 500:])
    assert a.host.w[0, 4] == 1.01


def test_spec_check_rejects_bad_models():
    good = GOLDENS["models"]["resources/example.clj"]
    def broken(path, value):
        import copy
        d = copy.deepcopy(good)
        t = d
        for k in path[:-1]:
            t = t[k]
        if value is None:
            del t[path[-1]]
        else:
            t[path[-1]] = value
        return as_model(d)
    for path, value in (([":line", ":m2", ":mu"], 0), ([":line", ":m2", ":lambda"], -0.1), ([":line", ":b1", ":N"], 0),
                        ([":line", ":m1", ":discipline"], ":XYZ"), ([":params", ":run-to-time"], None),
                        ([":jobmix", ":jobType1", ":portion"], 0.0), ([":jobmix", ":jobType2", ":w", ":m3"], -1.0),
                        ([":entry-point"], ":m2"), ([":topology"], [":m1", ":b1", ":m2"])):
        with pytest.raises(model.ModelError):
            model.preprocess_model(broken(path, value))


# ---- substream arithmetic: third, independent implementation ----------------------------------------

def test_numpy_philox_known_answers_and_oracle_agreement(oracle):
    kat = [([0, 0, 0, 0], [0, 0], [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]),
           ([0xffffffff] * 4, [0xffffffff] * 2, [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]),
           ([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0],
            [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1])]          # Random123 kat_vectors, philox4x32 10
    for ctr, key, want in kat:
        assert [int(x[0]) for x in philox.philox4x32_10(*ctr, *key)] == want
        assert oracle.philox(ctr, key) == want
    ids = np.array([0, 1, 77, 2 ** 33 + 5], np.uint64)
    u = philox.uniform53(SEED, ids, 9, 1234)
    for g, x in zip(ids, u):
        assert oracle.lib().mjpo_uniform(SEED, int(g), 9, 1234) == x


def test_deterministic_log_exp_accuracy(oracle):
    rng = np.random.default_rng(0)
    l = oracle.lib()
    xs = rng.random(20000) * 10.0 ** rng.uniform(-30, 0, 20000)
    assert max(abs(l.mjpo_log(float(x)) - math.log(x)) / abs(math.log(x)) for x in xs) < 4e-16
    ys = -rng.random(20000) * 40
    assert max(abs(l.mjpo_exp(float(y)) - math.exp(y)) / math.exp(y) for y in ys) < 4e-16
    assert l.mjpo_exp(-41.0) == 0.0 and l.mjpo_exp(0.0) == 1.0


# ---- results -------------------------------------------------------------------------------------------

def test_calc_bneck_goldens():
    g = GOLDENS["G6"]
    for tp, exp in ((g["tp1"], g["tp1_expected"]), (g["tp2"], g["tp2_expected"])):
        assert model.calc_bneck(tp, g["machines"])[":bottleneck-machine"] == exp


def test_calc_basics_matches_oracle_restatement(oracle):
    cl = model.preprocess_model(as_model(GOLDENS["models"]["data/quick.clj"]))
    r = oracle.run(cl.host, 3, seed=SEED)
    for g in range(3):
        a, b = model.calc_basics(cl, r.stats, g), oracle.calc_basics(cl.host, r.stats, g)
        assert a[":TP"] == b["TP"] and a[":number-of-jobs"] == b["njobs"]
        assert a[":observed-residence-time"] == b["residence"]
        assert list(a[":blocked"].values()) == b["blocked"] and list(a[":starved"].values()) == b["starved"]
        assert list(a[":avg-buffer-occupancy"].values()) == b["occupancy"]
        res = model.postprocess_model(cl, r.stats, g, g, 0.1)
        assert res[":status"] == ":normal-end" and res[":bneck"] == res[":bottleneck-machine"] in cl.machines


def test_calc_residence_time_and_diag_log_buf(oracle):
    """core.clj:663-700 (dormant in the reference) and :diag-log-buf? (core.clj:585,713-714; util/log.clj:134-135)."""
    d = dict(GOLDENS["models"]["data/quick.clj"])
    d[":report"] = {":log?": True, ":max-lines": 50, ":diag-log-buf?": True}
    cl = model.preprocess_model(as_model(d))
    r = oracle.run(cl.host, 2, seed=SEED, log_capacity=400)
    res = model.postprocess_model(cl, r.stats, 1, 1, 0.1, log=r.log)
    buf = res[":diag-log-buf"]
    want = scada.pull_data(cl, r.log, None, instance=1)
    assert len(buf) >= 50 and all(":line" not in m for m in buf)
    assert [dict(m, **{":line": w[":line"]}) for m, w in zip(buf, want)] == want
    # hand evaluation for a 2-machine, 1-type line: eff = mu/(lam+mu); w/W; wait = (0.5 + occ b1) * virt m2 * (1+bl m2)
    d2 = dict(GOLDENS["models"]["data/submodel-1.clj"])
    cl2 = model.preprocess_model(as_model(d2), n_inst=1)
    h = cl2.host
    fake = {":blocked": {cl2.machines[0]: 0.25, cl2.machines[1]: 0.0}, ":avg-buffer-occupancy": {cl2.buffers[0]: 1.5}}
    got = model.calc_residence_time(fake, cl2)[":computed-residence-time"]
    eff = [h.mu[i] / (h.lam[i] + h.mu[i]) for i in range(2)]
    req = [h.w[0, i] / h.W[i] for i in range(2)]
    virt2 = h.portion[0] * req[1] / eff[1]
    expect = req[0] / eff[0] * 1.25 + req[1] / eff[1] * 1.0 + (0.5 + 1.5) * virt2 * 1.0
    assert list(got) == cl2.jobtypes and got[cl2.jobtypes[0]] == pytest.approx(expect, rel=1e-14)


# ---- SCADA text ------------------------------------------------------------------------------------------

def _render(oracle, gid, warm_up=0):
    d = dict(GOLDENS[gid]["model"])
    d[":report"] = {":log?": True, ":max-lines": math.inf}
    d[":params"] = {":warm-up-time": warm_up, ":run-to-time": 1000}
    cl = model.preprocess_model(as_model(d))
    r = oracle.run(cl.host, 1, mode=abi.RNG_REPLAY, seed=SEED, log_capacity=2000)
    return cl, r, scada.render_lines(cl, r.log)


def test_scada_text_bbs_matches_log_test_verbatim(oracle):
    """log_test.clj:62-68 quotes six printed lines of the BBS cycle; with the test's own model
    (warm-up 0) they are reproduced character for character, :line numbers 14-19 included."""
    cl, r, text = _render(oracle, "G9")
    want = GOLDENS["G9"]["scada_text"]
    first = text.index(want[0])
    assert text[first:first + 6] == want
    msgs = scada.pull_data(cl, r.log, 500)
    assert len(msgs) == 500 and msgs[0][":line"] == 0 and list(msgs[0])[:4] == [":clk", ":act", ":m", ":mjpact"]


def test_scada_text_bas_matches_log_test(oracle):
    """log_test.clj:45-51 (BAS cycle at clk 5.6 / 6.8).  Same content at warm-up 0 but the quoted :line
    numbers (5-10) belong to a run that started printing at the 4.8 tick: with :warm-up-time 4.0 the six
    lines are reproduced verbatim."""
    import re
    want = GOLDENS["G8"]["scada_text"]
    strip = lambda t: re.sub(r" :line \d+", "", t)
    _, _, text0 = _render(oracle, "G8", warm_up=0)
    stripped = [strip(t) for t in text0]
    first = stripped.index(strip(want[0]))
    assert stripped[first:first + 6] == [strip(t) for t in want]
    _, _, text4 = _render(oracle, "G8", warm_up=4.0)
    first = text4.index(want[0])
    assert text4[first:first + 6] == want


def test_java_double_printing():
    for x, s in ((6.8, "6.8"), (1.6, "1.6"), (0.0, "0.0"), (100.0, "100.0"), (1.0e7, "1.0E7"), (12345678.5, "1.23456785E7"),
                 (0.001, "0.001"), (0.0001, "1.0E-4"), (52.8918, "52.8918")):
        assert scada.java_double(x) == s


# ---- end to end on the device ---------------------------------------------------------------------------

@pytest.mark.gpu
def test_main_loop_example_on_device(engine):
    d = dict(GOLDENS["models"]["data/quick.clj"])
    res = model.main_loop(as_model(d), handle=engine, seed=SEED)
    assert res[":status"] == ":normal-end" and 0.6 < res[":TP"] < 0.85
    assert set(res[":blocked"]) == {":m1", ":m2", ":m3", ":m4", ":m5"} and res[":blocked"][":m5"] == 0.0
    assert res[":starved"][":m1"] == 0.0 and res[":bottleneck-machine"] in res[":blocked"]
    assert set(res[":avg-buffer-occupancy"]) == {":b1", ":b2", ":b3", ":b4"}


@pytest.mark.gpu
def test_main_loop_multi_with_bounds_and_log(engine, oracle):
    d = dict(GOLDENS["models"]["data/example-10.clj"])
    d[":params"] = {":warm-up-time": 100, ":run-to-time": 600}
    d[":number-of-simulations"] = 6
    d[":report"] = {":log?": True, ":max-lines": 40}
    out = io.StringIO()
    res = model.main_loop(as_model(d), handle=engine, seed=SEED, out_stream=out)
    assert len(res) == 6 and len({r[":TP"] for r in res}) > 1
    text = out.getvalue()
    assert text.startswith("#_{") and text.count("\n{:clk") >= 6 * 40 and text.count("#_{") == 7
    # the same compiled line through the oracle gives the same numbers
    cl = model.preprocess_model(as_model(d), seed=SEED)
    want = oracle.run(cl.host, 6, seed=SEED)
    for g, r in enumerate(res):
        assert r[":number-of-jobs"] == int(want.stats.njobs[g])


@pytest.mark.gpu
@pytest.mark.parametrize("gid", ["G8", "G9"])
def test_log_test_clj_through_the_streaming_api(engine, gid):
    """util/log_test.clj reliable-BAS-pattern / reliable-BBS-pattern, as written there: main-loop on a model with
    :report {:continuous? true}, (pull-data! model 500), drop the warm-up, partition by 6, compare acts and spans."""
    g = GOLDENS[gid]
    cm = model.main_loop(as_model(g["model"]), handle=engine, seed=SEED)
    assert isinstance(cm, model.ContinuousModel)
    msgs = model.pull_data(cm, g["pulled"])
    assert len(msgs) == g["pulled"] and [m[":line"] for m in msgs] == list(range(g["pulled"]))
    log = [m for m in msgs if m[":clk"] >= g["from_clk"]]
    groups = [log[i:i + 6] for i in range(0, len(log) - len(log) % 6, 6)]
    assert len(groups) > 50
    for grp in groups:
        assert [m[":act"] for m in grp] == g["cycle"]
        assert abs((grp[-1][":clk"] - grp[0][":clk"]) - g["span"]) < g["tol"]
    more = model.pull_data(cm, 1000)                       # keeps going past :run-to-time 100 (core.clj:801-818)
    assert more[-1][":clk"] > 100.0 and more[0][":line"] == g["pulled"]


@pytest.mark.gpu
def test_streaming_equals_one_shot_capture(engine, oracle):
    d = dict(GOLDENS["models"]["data/quick.clj"])
    d[":report"] = {":continuous?": True, ":up&down?": True}
    cm = model.main_loop(as_model(d), handle=engine, seed=SEED, mode=abi.RNG_REPLAY)
    msgs = model.pull_data(cm, 3000)
    cl = cm.cl
    line = cl.host
    line.run_to = float(msgs[-1][":clk"]) + 50.0
    want = oracle.run(line, 1, mode=abi.RNG_REPLAY, seed=SEED, log_capacity=100_000)
    ref = scada.pull_data(cl, want.log, 3000)
    assert msgs == ref


def test_recorded_outputs_pin_hash_map_order_and_pprint_layout():
    """G11: the reference's recorded result blocks (data/*-out.clj), as text.  They were printed from 11-entry hash
    maps: (a) their key order is the iteration order cljhash.py computes for those keywords -- the same order rule
    the kernel and the oracle apply to ticks with more than 8 machines, to :jobmix beyond 8 job types and to
    calc-bneck beyond 8 machines; (b) re-printing the parsed block with scada.pprint_edn gives back the text byte
    for byte (clojure.pprint layout, Double/toString digits)."""
    from mjpdes_b200 import cljhash, edn, scada
    printed = GOLDENS["G11"]["printed"]
    assert len(printed) == 6
    for text in printed:
        form = edn.read_string(text)
        keys = list(form.keys())
        assert len(keys) == 11
        assert cljhash.map_order(sorted(keys)) == keys                     # (a)
        assert scada.pprint_edn(form) == text                              # (b)


def test_result_block_and_model_block_follow_the_reference_print_order():
    from mjpdes_b200 import scada
    res = {":runtime": 10.823, ":TP": 0.7300555944442749, ":observed-residence-time": 12.205651350281991,
           ":number-of-jobs": 13141, ":blocked": {":m1": 0.5, ":m2": 0.0}, ":starved": {":m1": 0.0, ":m2": 0.25},
           ":avg-buffer-occupancy": {":b1": 1.5}, ":bottleneck-machine": ":m1", ":process-id": 1, ":jobmix": {},
           ":status": ":normal-end", ":average-buffer-occupancy": {":b1": 1.5}, ":bneck": ":m1"}
    text = scada.result_block(res)
    assert text.startswith("#_{:TP 0.7300556,\n :number-of-jobs 13141,\n :avg-buffer-occupancy {:b1 1.5},\n :status :normal-end,")
    assert ":jobmix" not in text and ":bneck" not in text and text.endswith(":process-id 1}")
    # *print-length* 10 (core.clj:750,773): longer collections are cut
    long = scada.pprint_edn({":topology": [f":m{i}" for i in range(1, 14)]}, print_length=10)
    assert long == "{:topology [:m1 :m2 :m3 :m4 :m5 :m6 :m7 :m8 :m9 :m10 ...]}"
    # the model block: record fields first, in defrecord order and nil when unset (core.clj:38-66); the 9-entry :line
    # map in hash order; machines reduced to a plain map by pretty-model's dissoc (util/log.clj:293-303)
    m = as_model(GOLDENS["models"]["resources/example.clj"])
    block = scada.pprint_edn(scada.pretty_model(m), print_length=10)
    lines = block.split("\n")
    assert lines[0] == "{:line" and lines[1] == " {:m5 {:lambda 0.1, :mu 0.9, :W 1.2},"
    assert lines[-1] == " :print nil}" and " :topology [:m1 :b1 :m2 :b2 :m3 :b3 :m4 :b4 :m5]," in lines
    assert [l.split()[0] for l in lines if l.startswith(" :")] == [":entry-point", ":jobmix", ":params", ":topology", ":print"]
