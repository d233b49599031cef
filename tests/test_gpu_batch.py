"""GPU parity tests of the device-resolved RUN of proposals (DeviceStore.run_batch -> frmc_run_batch ->
batch_kernel): Engine.__on_runtime_step_try_move (Engine.py:3302-3338) for many proposals per launch.

Bars (all bit-exact):
* against the sequential device path driven by the same rule on the host (which tests/test_gpu_store.py and
  tests/test_golden_constraints.py pin to the reference): chi^2 of every proposal, every decision, the random
  numbers consumed, the final total standard error, the ordered data arrays, model totals, coordinates;
* against the oracle: the final histograms recomputed from scratch on the final coordinates (symmetrised);
* against the golden trajectories of the unmodified reference constraint classes (NiTi, THF, SiOx, synthetic),
  with the recorded accept/reject sequence forced through the pre-drawn random numbers.
"""
import numpy as np
import pytest

import cases as C
import test_gpu_store as TS
import test_golden_constraints as TG

pytestmark = pytest.mark.gpu
F32 = np.float32


def host_total(chi2, var2):
    """np.sum([SD / c.varianceSquared ...]) in float32 (Engine.py:3024-3029)"""
    return np.sum([F32(c) / F32(v) for c, v in zip(chi2, var2)], dtype=F32)


def sequential_run(store, proposals, total, rand, tolerance, var2):
    """the engine's rule on the host around one device launch per proposal"""
    chis, decs, used = [], [], 0
    total = F32(total)
    for idx, moved in proposals:
        chi2 = store.propose(idx, moved)
        nt = host_total(chi2, var2)
        dec = 1
        if nt > total:
            dec = 0 if rand[used] > F32(tolerance) else 2
            used += 1
        (store.accept if dec else store.reject)()
        if dec:
            total = nt
        chis.append(chi2.copy()); decs.append(dec)
    return np.array(chis, F32), np.array(decs, np.int32), total, used


def make_proposals(case, rng, n, sigma, repeat_every=0):
    box = case["boxCoords"]
    props = []
    for j in range(n):
        if repeat_every and j % repeat_every == repeat_every - 1 and props:
            idx = props[-1][0].copy()                        # the same atoms again: a conflict when the first one is accepted
        else:
            idx = C.group_for(case, rng)
        moved = (box[idx] + rng.normal(0, sigma, (idx.shape[0], 3)).astype(F32)).astype(F32)
        props.append((idx, moved))
    return props


def flatten(props):
    idx = np.concatenate([p[0] for p in props]).astype(np.int32)
    moved = np.concatenate([p[1] for p in props]).astype(F32)
    sizes = np.array([p[0].shape[0] for p in props], np.int32)
    return idx, moved, sizes


def compare_stores(a, b, n_models, n_grids=1):
    for g in range(n_grids):
        ai, ae = a.export_data(g)
        bi, be = b.export_data(g)
        assert np.array_equal(ai, bi) and np.array_equal(ae, be), "ordered data arrays differ (grid %d)" % g
    for m in range(n_models):
        assert np.array_equal(a.export_total(m), b.export_total(m)), "committed total of model %d differs" % m
    assert np.array_equal(a.get_coords(), b.get_coords())
    assert np.array_equal(a.committed_chi2(), b.committed_chi2())
    assert a.edge_overflow == b.edge_overflow


CONFIGS = [
    # case, kinds, n proposals, sigma, tolerance, repeat_every, model kwargs
    ("ortho_atomic", ["PDF", "SQ"], 90, 0.02, 0.0, 0, {}),
    ("ortho_atomic", ["PDF", "SQ"], 70, 0.02, 0.35, 7, {}),
    ("tri_molecular", ["PCF"], 40, 0.02, 0.2, 0, dict(with_weights=True)),
    ("tri_unwrapped", ["RSQ", "PDF"], 60, 0.05, 0.1, 5, dict(scale=0.93, with_shape=True)),
    ("ortho_unwrapped", ["PDF"], 64, 0.2, 0.5, 0, {}),
    ("ibc_nanoparticle", ["PDF", "PCF"], 50, 0.3, 0.15, 6, dict(with_shape=True, scale=1.07)),
    ("coincident_empty_class", ["PDF", "SQ"], 45, 0.02, 0.3, 4, {}),
    ("cfg4_small", ["PDF", "SQ"], 100, 0.02, 0.05, 0, {}),
    ("tiny_13", ["PDF"], 20, 0.05, 0.5, 3, {}),
]


@pytest.mark.parametrize("name,kinds,n,sigma,tol,repeat,mkw", CONFIGS,
                         ids=["%s-%s-tol%g" % (c[0], "+".join(c[1]), c[4]) for c in CONFIGS])
def test_batch_equals_sequential_device_path(name, kinds, n, sigma, tol, repeat, mkw, orc):
    case = TS.CASES[name]
    seq, _ = TS._build(case, kinds, np.random.default_rng(11), **mkw)
    bat, _ = TS._build(case, kinds, np.random.default_rng(11), **mkw)
    nm = len(kinds)
    var2 = np.array([1.0, 0.37, 2.5][:nm], F32)
    rng = np.random.default_rng(5)
    c0 = seq.compute_data()
    c1 = bat.compute_data()
    assert np.array_equal(c0, c1)
    total0 = host_total(c0, var2)
    props = make_proposals(case, rng, n, sigma, repeat)
    rand = rng.random(n).astype(F32)
    chis, decs, total, used = sequential_run(seq, props, total0, rand, tol, var2)
    idx, moved, sizes = flatten(props)
    out = bat.run_batch(idx, moved, total0, rand, tolerance=tol, group_sizes=sizes, variance_squared=var2)
    assert np.array_equal(out["decisions"], decs), "decisions differ: %s vs %s" % (out["decisions"], decs)
    assert np.array_equal(out["chi2"], chis), "chi2 of the proposals differ"
    assert F32(out["total"]) == F32(total) and out["rand_used"] == used
    assert 0 < int((decs > 0).sum()) < n, "degenerate sequence (%d accepted of %d)" % (int((decs > 0).sum()), n)
    compare_stores(seq, bat, nm)
    launches, rounds, resolved = bat.batch_stats()
    assert resolved == n and launches >= 1 and rounds >= 1
    # from scratch on the final coordinates (symmetrised: SURVEY 3.3)
    kw = TS._hist_kw(case)
    fi, fe = orc.full_pairs_histograms_coords(boxCoords=bat.get_coords(), moleculeIndex=case["moleculeIndex"],
                                              elementIndex=case["elementIndex"], **kw)
    gi, ge = bat.export_data(0)
    sym = lambda h: h + h.transpose(1, 0, 2)
    assert np.array_equal(sym(gi), sym(fi)) and np.array_equal(sym(ge), sym(fe))
    # the two paths keep mixing: a few single steps on both, then another run on both
    more = make_proposals(dict(case, boxCoords=bat.get_coords()), rng, 6, sigma)
    r2 = rng.random(6).astype(F32)
    a = sequential_run(seq, more, total, r2, tol, var2)
    b = sequential_run(bat, more, total, r2, tol, var2)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    again = make_proposals(dict(case, boxCoords=bat.get_coords()), rng, 12, sigma)
    r3 = rng.random(12).astype(F32)
    i3, m3, s3 = flatten(again)
    oa = seq.run_batch(i3, m3, a[2], r3, tolerance=tol, group_sizes=s3, variance_squared=var2)
    ob = bat.run_batch(i3, m3, b[2], r3, tolerance=tol, group_sizes=s3, variance_squared=var2)
    assert np.array_equal(oa["chi2"], ob["chi2"]) and np.array_equal(oa["decisions"], ob["decisions"])
    compare_stores(seq, bat, nm)
    seq.close(); bat.close()


def test_long_run_spans_several_launches():
    """4 500 proposals = 141+ batches: more than one launch of the multi-batch kernel (128 batches each), with a repeated
    atom every 97 proposals (a conflict ends a launch in the middle and the host cuts the rest again); chi^2 of every
    proposal, every decision and the final state equal the sequential device path"""
    case = TS.CASES["ortho_atomic"]
    kinds = ["PDF", "SQ"]
    seq, _ = TS._build(case, kinds, np.random.default_rng(3))
    bat, _ = TS._build(case, kinds, np.random.default_rng(3))
    var2 = np.array([1.0, 0.5], F32)
    rng = np.random.default_rng(17)
    total0 = host_total(seq.compute_data(), var2)
    bat.compute_data()
    n = 4500
    props = make_proposals(case, rng, n, 0.02, 97)
    rand = rng.random(n).astype(F32)
    chis, decs, total, used = sequential_run(seq, props, total0, rand, 0.2, var2)
    idx, moved, sizes = flatten(props)
    l0 = bat.batch_stats()[0]
    out = bat.run_batch(idx, moved, total0, rand, tolerance=0.2, group_sizes=sizes, variance_squared=var2)
    assert np.array_equal(out["decisions"], decs) and np.array_equal(out["chi2"], chis)
    assert F32(out["total"]) == F32(total) and out["rand_used"] == used
    assert 0 < int((decs > 0).sum()) < n
    compare_stores(seq, bat, 2)
    launches = bat.batch_stats()[0] - l0
    assert 2 <= launches < n // 32, "expected a few multi-batch launches, saw %d" % launches
    seq.close(); bat.close()


@pytest.mark.parametrize("kind", ["ortho", "tri_unwrapped", "non_periodic"])
def test_batch_on_large_sparse_systems(kind, orc):
    """maxDistance far below the cell size: the delta pass skips most (sub-block, moved atom) pairs, most proposals
    do not interact, and the rounds resolve several acceptances at once on speculative evaluations (committed state
    + assumed-accepted deltas).  Proposals next to an earlier one (pair corrections), repeated atoms (conflicts),
    long jumps across the seam and molecule moves are mixed in.  Bars: bit-exact against the sequential device path,
    against the same run with culling switched off, and against the oracle's histogram of the final coordinates."""
    from fullrmc_b200 import _lib as fullrmc_b200
    case = TS._large_sparse_case(kind)
    box0 = case["boxCoords"]
    real = box0.astype(np.float64) @ np.asarray(case["basis"], np.float64)
    rng = np.random.default_rng(77)
    n = 150
    span = 1.0 if case["isPBC"] else 100.0
    props = []
    for j in range(n):
        if j % 9 == 8:
            idx = props[-1][0].copy()                                      # the same atoms again (conflict if accepted)
        elif j % 3 == 1:
            d = np.linalg.norm(real - real[props[-1][0][0]], axis=1)       # an atom a few A from the previous proposal's
            idx = np.array([int(np.argsort(d)[1 + j % 5])], np.int32)
        else:
            idx = C.group_for(case, rng)
        jump = (0.45 if j % 7 == 0 else 0.004) * span
        moved = (box0[idx] + rng.normal(0, jump, (idx.shape[0], 3)).astype(F32)).astype(F32)
        props.append((idx, moved))
    rand = rng.random(n).astype(F32)
    idx, moved, sizes = flatten(props)
    var2 = np.array([1.0, 0.6], F32)
    results = []
    for mode in ("sequential", "batch", "batch_noculling"):
        st, _ = TS._build(case, ["PDF", "SQ"], np.random.default_rng(11))
        total0 = host_total(st.compute_data(), var2)
        if mode == "sequential":
            chis, decs, total, used = sequential_run(st, props, total0, rand, 0.3, var2)
        else:
            previous = fullrmc_b200.set_block_culling(mode == "batch")
            try:
                out = st.run_batch(idx, moved, total0, rand, tolerance=0.3, group_sizes=sizes, variance_squared=var2)
            finally:
                fullrmc_b200.set_block_culling(previous)
            chis, decs, total, used = out["chi2"], out["decisions"], out["total"], out["rand_used"]
            launches, rounds, resolved = st.batch_stats()
            assert resolved == n and rounds < n, "no round resolved more than one proposal (%d rounds)" % rounds
        results.append((st, chis, decs, F32(total), used))
    s0, chis, decs, total, used = results[0]
    assert 10 < int((decs > 0).sum()) < n - 10
    for st, c, d, t, u in results[1:]:
        assert np.array_equal(d, decs), "decisions differ"
        assert np.array_equal(c, chis), "chi2 of the proposals differ"
        assert t == total and u == used
        compare_stores(s0, st, 2)
    kw = TS._hist_kw(case)
    fi, fe = orc.full_pairs_histograms_coords(boxCoords=s0.get_coords(), moleculeIndex=case["moleculeIndex"],
                                              elementIndex=case["elementIndex"], ncores=orc.max_threads(), **kw)
    gi, ge = results[1][0].export_data(0)
    sym = lambda h: h + h.transpose(1, 0, 2)
    assert np.array_equal(sym(gi), sym(fi)) and np.array_equal(sym(ge), sym(fe))
    for r in results:
        r[0].close()


@pytest.mark.parametrize("tol", [1.0, 0.0])
def test_batch_deep_chains(tol, orc):
    """single-atom proposals far apart in a large sparse system: nothing interacts, so the rounds' chains of predicted
    outcomes run as deep as the plan allows.  tolerance 1 accepts every proposal (every node of a chain assumes all the
    proposals before it, the limit of assumed acceptances and the pending commits are always at their maximum);
    tolerance 0 is the benchmark's rule.  Bit-exact against the sequential device path."""
    case = TS._large_sparse_case("ortho")
    rng = np.random.default_rng(123)
    n = 96
    atoms = rng.choice(case["boxCoords"].shape[0], n, replace=False).astype(np.int32)
    props = [(atoms[j:j + 1], (case["boxCoords"][atoms[j:j + 1]] + rng.normal(0, 0.003, (1, 3)).astype(F32)).astype(F32)) for j in range(n)]
    rand = rng.random(n).astype(F32)
    idx, moved, sizes = flatten(props)
    var2 = np.array([1.0, 0.8], F32)
    seq, _ = TS._build(case, ["PDF", "SQ"], np.random.default_rng(11))
    bat, _ = TS._build(case, ["PDF", "SQ"], np.random.default_rng(11))
    total0 = host_total(seq.compute_data(), var2)
    bat.compute_data()
    chis, decs, total, used = sequential_run(seq, props, total0, rand, tol, var2)
    out = bat.run_batch(idx, moved, total0, rand, tolerance=tol, group_sizes=sizes, variance_squared=var2)
    assert np.array_equal(out["decisions"], decs) and np.array_equal(out["chi2"], chis)
    assert F32(out["total"]) == F32(total) and out["rand_used"] == used
    if tol == 1.0:
        assert int((decs > 0).sum()) == n
    compare_stores(seq, bat, 2)
    launches, rounds, resolved = bat.batch_stats()
    assert resolved == n and rounds <= n // 2, "the chains resolved fewer than two proposals per round (%d rounds)" % rounds
    seq.close(); bat.close()


def test_batch_two_grids(orc):
    """PDF and S(Q) on different r-grids (the NiTi arrangement): both grids' deltas come from the same pass"""
    from fullrmc_b200.model import ModelSpec
    from fullrmc_b200.store import DeviceStore
    from oracle import epilogue as ep
    case = TS.CASES["ortho_atomic"]
    els, n_per, wdict, volume, rho0 = TS._system_meta(case)
    rng = np.random.default_rng(3)

    def build():
        st = DeviceStore(case["boxCoords"], case["basis"], True, case["moleculeIndex"], case["elementIndex"], case["numberOfElements"])
        r = np.random.default_rng(17)
        for (rmin, b, hs, kind) in ((0.0, 0.02, 650, "PDF"), (0.3, 0.2, 60, "RSQ")):
            edges = (rmin + b * np.arange(hs + 1, dtype=np.float64)).astype(F32)
            centers = ((edges[:-1] + edges[1:]) / F32(2.)).astype(F32)
            g = st.add_grid(edges[0], edges[-1], F32(b), hs)
            q = np.linspace(0.4, 12.0, 75).astype(F32) if kind == "RSQ" else None
            n_out = hs if q is None else q.shape[0]
            st.add_model(g, ModelSpec(kind, els, n_per, wdict, volume, rho0, centers, ep.shell_arrays_from_edges(edges),
                                      r.normal(0, 0.4, n_out).astype(F32), q_values=q))
        return st
    seq, bat = build(), build()
    var2 = np.array([1.0, 1.0], F32)
    total0 = host_total(seq.compute_data(), var2)
    bat.compute_data()
    props = make_proposals(case, rng, 75, 0.03, repeat_every=9)
    rand = rng.random(75).astype(F32)
    chis, decs, total, used = sequential_run(seq, props, total0, rand, 0.2, var2)
    idx, moved, sizes = flatten(props)
    out = bat.run_batch(idx, moved, total0, rand, tolerance=0.2, group_sizes=sizes, variance_squared=var2)
    assert np.array_equal(out["decisions"], decs) and np.array_equal(out["chi2"], chis)
    assert F32(out["total"]) == F32(total) and out["rand_used"] == used
    compare_stores(seq, bat, 2, n_grids=2)
    seq.close(); bat.close()


def forced_random_numbers(g):
    """rand / tolerance that make the engine's rule reproduce a golden trajectory's recorded decisions: a worse
    proposal consumes one number, 0 (<= tolerance: accepted) or 1 (rejected); better ones are always accepted,
    which is also what the recording did (tests/gen_golden_constraints.py)"""
    chi = g["steps/chi2_after"]
    acc = g["steps/accepted"]
    nc = int(g["n_constraints"])
    total = np.sum([F32(x) for x in g["start_stdErr"][:nc]], dtype=F32)
    rand = []
    for s in range(chi.shape[0]):
        nt = np.sum([F32(x) for x in chi[s, :nc]], dtype=F32)
        if nt > total:
            rand.append(0.0 if acc[s] else 1.0)
        elif not acc[s]:
            return None                                        # the recording rejected a better move: cannot be forced
        if acc[s]:
            total = nt
    rand += [1.0] * (chi.shape[0] - len(rand))
    return np.array(rand, F32), np.sum([F32(x) for x in g["start_stdErr"][:nc]], dtype=F32)


@pytest.mark.parametrize("name", ["niti", "thf", "siox", "synth", "niti_sf", "synth_sf"])
def test_batch_reproduces_reference_trajectories(name, golden_dir):
    """*_sf: scale-factor refit schedules run inside the batch kernel (a node refits when the engine's accepted count at
    that node is a multiple of the frequency; an accepted refit changes the scale factor of everything behind it)"""
    import fullrmc_b200
    from fullrmc_b200.constraints import DeviceBackend, make_device_constraint
    g = TG._load(golden_dir, name)
    forced = forced_random_numbers(g)
    assert forced is not None
    rand, total0 = forced
    previous = fullrmc_b200.set_edge_spill(True)
    try:
        elements, n_per = TG._system(g)
        backend = DeviceBackend(g["boxCoords"], g["basis"], bool(g["isPBC"]), g["moleculeIndex"], g["elementIndex"], elements,
                                n_per, g["volume"], g["numberDensity"])
        nc = int(g["n_constraints"])
        cons = []
        for ci in range(nc):
            d = TG._constraint_desc(g, ci)
            cons.append((d, make_device_constraint(backend, d["kind"], d["experimental"], d["minDistance"], d["maxDistance"], d["bin"],
                                                   int(d["histSize"]), d["shellCenters"], d["shellVolumes"], d["weighting"],
                                                   dataWeights=d["dataWeights"], shapeArray=d["shapeArray"],
                                                   scaleFactor=float(d["scaleFactor"]),
                                                   qValues=d.get("qValues") if d["kind"] in ("SQ", "RSQ") else None,
                                                   adjustScaleFactor=d["adjust"])))
        for ci, (d, c) in enumerate(cons):
            _, err = c.compute_data()
            assert F32(err) == F32(g["start_stdErr"][ci])
        steps = g["steps/idx"].shape[0]
        ks = g["steps/k"][:steps].astype(np.int32)
        idx = np.concatenate([g["steps/idx"][s, :ks[s]] for s in range(steps)]).astype(np.int32)
        moved = np.concatenate([g["steps/moved"][s, :ks[s]] for s in range(steps)]).astype(F32)
        out = backend.store.run_batch(idx, moved, total0, rand, tolerance=0.5, group_sizes=ks)
        assert np.array_equal(out["decisions"] > 0, g["steps/accepted"][:steps])
        assert np.array_equal(out["chi2"], g["steps/chi2_after"][:steps, :nc].astype(F32)), "chi2 differs from the reference classes"
        for ci, (d, c) in enumerate(cons):
            hi, he = backend.store.export_data(c._grid)
            assert np.array_equal(hi, d["final_intra"]) and np.array_equal(he, d["final_inter"])
            assert F32(backend.store.committed_chi2()[c._model]) == F32(d["final_stdErr"])
            if not d["adjust"][0]:
                assert np.array_equal(backend.store.export_total(c._model), d["final_total"])
            if "c%d/final_scaleFactor" % ci in g.files:
                assert F32(backend.store.get_scale(c._model)[0]) == F32(g["c%d/final_scaleFactor" % ci])
        assert np.array_equal(backend.store.get_coords(), g["final_boxCoords"])
        launches, _, _ = backend.store.batch_stats()
        assert launches >= 1
        backend.close()
    finally:
        fullrmc_b200.set_edge_spill(previous)


def test_batch_argument_errors():
    from fullrmc_b200.store import DeviceStore
    case = TS.CASES["tiny_13"]
    st, _ = TS._build(case, ["PDF"], np.random.default_rng(0))
    one = np.zeros(1, np.int32)
    with pytest.raises(RuntimeError):                          # no committed data yet
        st.run_batch(one, case["boxCoords"][:1], 1.0, np.zeros(1, F32))
    st.compute_data()
    with pytest.raises(ValueError):
        st.run_batch(np.array([99], np.int32), case["boxCoords"][:1], 1.0, np.zeros(1, F32))
    with pytest.raises(ValueError):
        st.run_batch(one, np.full((1, 3), np.nan, F32), 1.0, np.zeros(1, F32))
    with pytest.raises(ValueError):
        st.run_batch(one, case["boxCoords"][:1], 1.0, np.zeros(0, F32))
    st.propose(one, case["boxCoords"][:1])
    with pytest.raises(RuntimeError):                          # a proposal is staged
        st.run_batch(one, case["boxCoords"][:1], 1.0, np.zeros(1, F32))
    st.reject()
    st.close()
