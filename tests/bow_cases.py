"""Seeded bag-of-words cases shared by the CPU tests, tests/golden/make_golden.py and the GPU tests."""
import numpy as np

from orb_slam2_refactored_b200 import synth

# (name, vocabulary kwargs, features, levelsup, scoring, weighting)
CASES = [
    ('k10L4', dict(seed=1, k=10, L=4), 1500, 2, 0, 0),                             # the reference's shape, shallower: L1_NORM, TF_IDF
    ('k10L4_l4', dict(seed=1, k=10, L=4), 700, 4, 0, 0),                           # levelsup >= L: every node id is the root
    ('k6L5_pruned', dict(seed=2, k=6, L=5, prune=0.25, min_leaf_level=3), 2000, 2, 0, 0),   # ragged tree, leaves at 3 depths
    ('k3L6_tf_l2', dict(seed=3, k=3, L=6), 900, 4, 1, 1),                          # TF + L2_NORM
    ('k8L3_idf_dot', dict(seed=4, k=8, L=3), 1200, 1, 5, 2),                       # IDF + DOT_PRODUCT: addIfNotExist, no normalisation
    ('k8L3_tf_dot', dict(seed=4, k=8, L=3), 1200, 1, 5, 1),                        # TF without normalisation: divided by the word count
    ('k20L2_binary_chi', dict(seed=5, k=20, L=2, stop=0.5), 3000, 0, 2, 3),        # BINARY + CHI_SQUARE (L1), half of the words stopped
    ('k9L3_scattered', dict(seed=8, k=9, L=3, prune=0.2, min_leaf_level=2, scatter=True), 1500, 1, 0, 0),   # siblings without consecutive ids
    ('k2L10_all_stopped', dict(seed=6, k=2, L=10, stop=1.0), 300, 4, 0, 0),        # every word stopped: both outputs empty
]
SMALL = ('k6L5_pruned', 'k8L3_idf_dot', 'k20L2_binary_chi')   # the cases whose reference outputs are committed as golden


def make(name):
    for c in CASES:
        if c[0] == name:
            _, kw, n, levelsup, scoring, weighting = c
            kw = dict(kw)
            scatter = kw.pop('scatter', False)
            voc = synth.vocabulary(**kw)
            if scatter:
                voc = synth.scatter_vocabulary(voc, kw['seed'])
            voc['scoring'], voc['weighting'] = scoring, weighting
            feats = synth.vocabulary_features(kw['seed'] + 100, voc, n)
            return voc, feats, levelsup
    raise KeyError(name)


def flatten(res):
    (wi, wv), (fn, fs, fi) = res[0:2] if len(res) == 2 else ((res[0], res[1]), res[2])
    return dict(word_ids=np.asarray(wi, np.int32), word_vals=np.asarray(wv, np.float64), fv_nodes=np.asarray(fn, np.uint32),
                fv_start=np.asarray(fs, np.int32), fv_items=np.asarray(fi, np.uint32))


def same(a, b):
    return all(a[k].tobytes() == b[k].tobytes() for k in ('word_ids', 'word_vals', 'fv_nodes', 'fv_start', 'fv_items'))
