#!/usr/bin/env python
"""Freeze the LIVE reference's rot_corr results at BASELINE configs[3] size (build container only).

TEST INFRASTRUCTURE.  Run:  python oracle/gen_golden_c4.py [--n-big 20000] [--skip-750]

Two fixtures on the 63-atom, six-rotor molecule of oracle/rotor_molecules.py (ensemble_tritbu63):

  * tritbu63_s11_750  — 750 structures, the reference UNMODIFIED (750 is the largest ensemble
    torsion_module.py:1056 lets through);
  * tritbu63_s13_20000 — 20 000 structures with the size guard lifted.  The reference refuses
    such an ensemble (returns an all-True mask), so this run executes the reference's own source
    of prune_conformers_rmsd_rot_corr with ONE literal changed (`len(structures) > 750` ->
    `len(structures) > 10**9`), compiled into the reference module's own namespace; every other
    line, and everything it calls, is the reference's.  Say "guard lifted" wherever it is quoted.

Both ran with the SURVEY A.6 stand-in for the un-vendored rmsd==1.4 (oracle/ref_harness.py).
Stored: mask (hex), digest, survivor count, wall time, the chemistry perception the GPU path
needs (torsions, angle sets, rotation masks, sub-graph node lists) and the returned (centred +
mutated) structures.
"""
import argparse
import copy
import inspect
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import ref_harness  # noqa: E402
import rotor_molecules as rm  # noqa: E402
from tscode_b200.synth import mask_digest  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n-big", type=int, default=20000)
    ap.add_argument("--skip-750", action="store_true")
    ap.add_argument("--skip-big", action="store_true")
    args = ap.parse_args()
    ref_harness.install(full=True)
    import networkx as nx
    import tscode.torsion_module as tm
    from tscode.graph_manipulations import graphize
    from tscode.utils import get_double_bonds_indices

    def perceive(ref, atomnos, graph):
        """set-up block of torsion_module.py:1026-1049 + the pair-independent quantities of :964-977, :301-325"""
        graph = copy.deepcopy(graph)
        for hb in tm._get_hydrogen_bonds(ref, atomnos, graph):
            graph.add_edge(*hb)
        tors = tm._get_torsions(graph, hydrogen_bonds=tm._get_hydrogen_bonds(ref, atomnos, graph),
                                double_bonds=get_double_bonds_indices(ref, atomnos), keepdummy=True)
        tors = [t for t in tors if not (tm._is_nondummy(t.i2, t.i3, graph) and tm._is_nondummy(t.i3, t.i2, graph))]
        tors = [t for t in tors if 1 not in [atomnos[i] for i in t.torsion]]
        angles = [t.get_angles() for t in tors]
        tors = [t.torsion if tm._is_nondummy(t.i2, t.i3, graph) else list(reversed(t.torsion)) for t in tors]
        masks, nodes = [], []
        for t in tors:
            for o in tors:
                if o is not t:
                    graph.remove_edge(o[1], o[2])
            nodes.append(sorted(i for i in [s for s in nx.connected_components(graph) if t[1] in s][0] if atomnos[i] != 1))
            for o in tors:
                if o is not t:
                    graph.add_edge(o[1], o[2])
            masks.append(tm._get_rotation_mask(graph, t))
        return [list(map(int, t)) for t in tors], [list(a) for a in angles], np.array(masks), nodes

    path = os.path.join(GOLD, "rotcorr_big.json")
    fixtures = json.load(open(path))["fixtures"] if os.path.exists(path) else {}

    def run(name, seed, N, fn, label):
        S, atomnos = rm.ensemble_tritbu63(seed, N)
        graph = graphize(S[0], atomnos)
        Sc0 = S[0] - S[0].mean(axis=0)
        tors, angles, masks, nodes = perceive(Sc0, atomnos, graph)
        t0 = time.perf_counter()
        out, mask = fn(S.copy(), atomnos, copy.deepcopy(graph), max_rmsd=0.25)
        dt = time.perf_counter() - t0
        fixtures[name] = dict(seed=seed, N=N, thr=0.25, torsions=tors, angles=angles, survivors=int(mask.sum()),
                              digest=mask_digest(mask), wall_s=round(dt, 1), reference=label,
                              mask_hex=np.packbits(mask.astype(np.uint8)).tobytes().hex())
        np.savez_compressed(os.path.join(GOLD, f"rotcorr_{name}.npz"), atomnos=atomnos, rot_masks=masks,
                            node_lists=np.array([np.isin(np.arange(len(atomnos)), n) for n in nodes]),
                            mask=mask, out=out)
        print("rotcorr_big:", name, {k: v for k, v in fixtures[name].items() if k != "mask_hex"}, flush=True)
        json.dump({"note": "rmsd==1.4 replaced by the Appendix A.6 stand-in; see oracle/gen_golden_c4.py",
                   "fixtures": fixtures}, open(path, "w"), indent=1)

    if not args.skip_750:
        run("tritbu63_s11_750", 11, 750, tm.prune_conformers_rmsd_rot_corr, "unmodified")
    if not args.skip_big:
        src = inspect.getsource(tm.prune_conformers_rmsd_rot_corr)
        assert src.count("len(structures) > 750") == 1
        ns = tm.__dict__
        saved = ns["prune_conformers_rmsd_rot_corr"]
        exec(compile(src.replace("len(structures) > 750", "len(structures) > 10**9"), "<guard-lifted reference>", "exec"), ns)
        lifted = ns["prune_conformers_rmsd_rot_corr"]
        ns["prune_conformers_rmsd_rot_corr"] = saved
        run(f"tritbu63_s13_{args.n_big}", 13, args.n_big, lifted,
            "guard lifted: `len(structures) > 750` -> `> 10**9` in the reference's own source, nothing else changed")


if __name__ == "__main__":
    main()
