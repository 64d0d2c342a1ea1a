import sys, importlib, numpy as np, time
sys.path.insert(0, '.')
from oracle.pyoracle import Oracle, Ref, slam_params
pkg = importlib.import_module('lc-crf-slam_b200')
synth = pkg.synth
o = Oracle()
ctx = pkg.Context(0)
rng = np.random.default_rng(0)
bad = 0
for d in (1, 2, 3, 4, 5, 6, 7):
    for N in (0, 1, 4, 5, 6, 7, 8, 257, 1000, 5003):
        f = (rng.normal(0, 3, (N, d))).astype(np.float32)
        if N > 8:
            f[::7] = np.round(f[::7]); f[::11] = np.round(f[::11] * 2) / 2
        lo = o.lattice(f)
        lg = pkg.Lattice(ctx, f)
        off, bary, nbr = lg.export()
        ok = lo['V'] == lg.V and np.array_equal(lo['offset'], off) and np.array_equal(lo['bary'].view(np.int32), bary.view(np.int32)) and np.array_equal(lo['nbr'], nbr)
        if not ok:
            bad += 1; print("LATTICE MISMATCH d", d, "N", N, "V", lo['V'], lg.V, "off", np.array_equal(lo['offset'], off), "bary", np.array_equal(lo['bary'].view(np.int32), bary.view(np.int32)), "nbr", np.array_equal(lo['nbr'], nbr))
        for L in (1, 2, 5):
            if N == 0: continue
            x = (rng.random((N, L)) * 3 - 0.5).astype(np.float32)
            yo = o.filter(lo, x); yg = lg.filter(x)
            err = np.abs(yo - yg).max() / max(np.abs(yo).max(), 1e-30)
            if not err < 1e-5:
                bad += 1; print("FILTER MISMATCH d", d, "N", N, "L", L, err)
        o.lattice_free(lo); lg.close()
print("lattice/filter mismatching cases:", bad)
prm_o = slam_params(**synth.SLAM_PARAMS); prm = pkg.SlamParams.make()
en = pkg.label_energies(2, prm.confidence)
for N in (2999, 3000, 3001, 3002, 100000):
    fr = synth.slam_frame(N, seed=N)
    lab = o.rough_classify(fr.observs, fr.error, fr.depth, prm_o)
    lab_g = ctx.rough_classify(fr.observs, fr.error, fr.depth, prm)
    Qo, mo, V = o.slam_crf(fr.observs, fr.error, fr.kp2d, lab, en, prm_o)
    crf = pkg.DenseCRF(ctx, N, 2)
    crf.setUnaryEnergyFromLabel(lab, energies=en)
    crf.addPairwiseEnergy(np.stack([fr.observs / np.float32(prm.stdev_beta), fr.error / np.float32(prm.stdev_alpha)], 1), prm.w1)
    crf.addPairwiseEnergy(fr.kp2d / np.float32(prm.point2d_stdev), prm.w2)
    t = time.time(); crf.inference(5, True); dt = time.time() - t
    Q = crf.getProbability(); m = crf.getMap()
    rel = np.abs(Q - Qo) / np.maximum(np.abs(Qo), 1e-30); rel[(Qo == 0) & (Q == 0)] = 0
    print(N, "V", V, crf.potts_vertices(0), crf.potts_vertices(1), "label diff", int((lab != lab_g).sum()), "max rel", rel.max(), "map diff", int((m != mo).sum()), "bit-equal frac", float((Q.view(np.int32) == Qo.view(np.int32)).mean()), "t %.2f ms" % (dt * 1e3))
    # frames path
    F = pkg.Frames(ctx, [N], prm, en)
    F.set_inputs(fr.observs, fr.error, fr.depth, fr.kp2d)
    F.run(); mp, pr = F.get_outputs()
    print("   frames: prob equal crf", np.array_equal(pr, Q), "map equal", np.array_equal(mp, m), F.get_debug()['V'])
    ctx.sync(); t = time.time()
    for _ in range(10): F.run()
    ctx.sync(); print("   frames run avg %.3f ms" % ((time.time() - t) * 100))
# unary
snap = synth.map_snapshot(5000, 64, seed=3, ragged=True)
ob, er, de = o.map_point_unary(snap); gob, ger, gde = ctx.map_point_unary(snap)
print("unary bit-equal:", np.array_equal(ob, gob), np.array_equal(er.view(np.int32), ger.view(np.int32)), np.array_equal(de.view(np.int32), gde.view(np.int32)), "err mean", er.mean())
print("launches", ctx.kernel_launches)
