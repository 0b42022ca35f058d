"""GPU parity tests of the bag-of-words transform, through the C ABI (orbx_vocabulary_*, orbx_bow_*): bit-exact word
ids, node ids, feature lists and FP64 word values against the oracle restatement, which tests/test_bow_oracle.py holds equal to the
reference's own DBoW2 text, and against the committed outputs of that text (tests/golden/bow.npz)."""
import os

import numpy as np
import pytest

import bow_cases
from orb_slam2_refactored_b200 import synth

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def _voc(orbx, voc):
    return orbx.ORBVocabulary().create(voc['k'], voc['L'], voc['parent'], voc['is_leaf'], voc['desc'], voc['weights'], voc['scoring'], voc['weighting'])


@pytest.mark.parametrize('name', [c[0] for c in bow_cases.CASES])
def test_transform_equals_oracle(orbx, oracle_port, name, tmp_path):
    voc, feats, levelsup = bow_cases.make(name)
    ov = oracle_port.vocabulary(arrays=voc)
    wi, wv, fv = ov.transform(feats, levelsup)
    want = bow_cases.flatten(((wi, wv), fv))
    gv = _voc(orbx, voc)
    bow, fvec, (fw, fnode) = gv.transform(feats, levelsup, per_feature=True)
    got = bow_cases.flatten((bow, fvec))
    ow, on = ov.descend(feats, levelsup)
    assert np.array_equal(fw, ow) and np.array_equal(fnode, on)          # the tree walk of every feature, stopped or not
    assert bow_cases.same(want, got)
    if name in bow_cases.SMALL:
        g = np.load(os.path.join(G, 'bow.npz'))
        for k in ('word_ids', 'word_vals', 'fv_nodes', 'fv_start', 'fv_items'):
            assert g[f'{name}_{k}'].tobytes() == got[k].tobytes(), k
    # the text loader (this fork's atoi tokenizer: fractional weights are truncated) builds the same tree
    path = str(tmp_path / 'voc.txt')
    synth.write_vocabulary_text(voc, path)
    tv = orbx.ORBVocabulary()
    assert tv.loadFromTextFile(path)
    assert tv.info()['nodes'] == len(voc['parent']) + 1 and tv.info()['words'] == int(voc['is_leaf'].sum())
    assert bow_cases.same(want, bow_cases.flatten(tv.transform(feats, levelsup)))
    # subsets and single features
    for n in (1, 2, 33, len(feats) // 2):
        wi, wv, fv = ov.transform(feats[:n], levelsup)
        assert bow_cases.same(bow_cases.flatten(((wi, wv), fv)), bow_cases.flatten(gv.transform(feats[:n], levelsup)))


def test_reference_shape_vocabulary(orbx, oracle_port):
    """k = 10, L = 6, levelsup = 4 (src/Frame.cc:213): 1.1 M nodes, 10^6 words; C1 / C4 feature counts."""
    voc = synth.vocabulary(7, 10, 6)
    assert len(voc['parent']) == 1111110
    ov = oracle_port.vocabulary(arrays=voc)
    gv = _voc(orbx, voc)
    assert gv.info() == dict(k=10, L=6, scoring=0, weighting=0, nodes=1111111, words=1000000)
    for seed, n in ((1, 1000), (2, 2005), (3, 8007), (4, 16384)):
        feats = synth.vocabulary_features(seed, voc, n)
        wi, wv, fv = ov.transform(feats, 4)
        want = bow_cases.flatten(((wi, wv), fv))
        got = bow_cases.flatten(gv.transform(feats, 4))
        assert bow_cases.same(want, got), n
        assert len(want['fv_nodes']) <= 100 and abs(want['word_vals'].sum() - 1.0) < 1e-9
    with pytest.raises(orbx.OrbxError):
        gv.transform(np.zeros((16385, 32), np.uint8), 4)
    (wi, wv), (fn, fs, fi) = gv.transform(np.zeros((0, 32), np.uint8), 4)
    assert len(wi) == 0 and len(fn) == 0 and list(fs) == [0]


def test_batch_device_after_extract(orbx, oracle_port):
    """Extract -> ComputeBoW with the descriptors staying on the device (src/System.cc:449 -> src/Tracking.cc:260)."""
    import torch
    voc = synth.vocabulary(8, 10, 4)
    ov = oracle_port.vocabulary(arrays=voc)
    gv = _voc(orbx, voc)
    ex = orbx.ORBextractor(nfeatures=1000)
    frames = np.stack([synth.image(s, 640, 480) for s in range(5)] + [np.full((480, 640), 90, np.uint8)])   # the last one has no keypoints
    d_frames = torch.from_numpy(frames).cuda()
    d_kps, d_desc, d_n = ex.extract_batch_device(d_frames)
    cap = d_kps.shape[1]
    res = gv.transform_batch_device(d_desc, d_n, len(frames), cap, 2, stream=ex.stream())
    ex.synchronize()
    word_ids, word_vals, fv_nodes, fv_start, fv_items, counts = [t.cpu().numpy() for t in res]
    n = d_n.cpu().numpy(); desc = d_desc.cpu().numpy()
    assert n[-1] == 0 and tuple(counts[-1]) == (0, 0)
    for f in range(len(frames)):
        wi, wv, fv = ov.transform(desc[f, :n[f]], 2)
        nw, nf = counts[f]
        assert nw == len(wi) and nf == len(fv[0])
        assert np.array_equal(word_ids[f, :nw], wi) and word_vals[f, :nw].tobytes() == wv.tobytes()
        assert np.array_equal(fv_nodes[f, :nf].view(np.uint32), fv[0]) and np.array_equal(fv_start[f, :nf + 1], fv[1])
        assert np.array_equal(fv_items[f, :fv_start[f, nf]].view(np.uint32), fv[2])


def test_score_pairs(orbx, oracle_port):
    voc = synth.vocabulary(9, 10, 4)
    ov = oracle_port.vocabulary(arrays=voc)
    gv = _voc(orbx, voc)
    vecs = []
    base = synth.vocabulary_features(50, voc, 1200)
    for s in range(12):
        f = base.copy()
        r = np.random.RandomState(s)
        k = r.randint(0, len(f), 100 * s)
        f[k] = synth.vocabulary_features(60 + s, voc, len(k))
        vecs.append(gv.transform(f, 2)[0])
    vecs.append((np.empty(0, np.int32), np.empty(0, np.float64)))
    pairs = [(a, b) for a in range(len(vecs)) for b in range(len(vecs))]
    got = gv.score_pairs(vecs, pairs)
    want = np.array([ov.score(vecs[a], vecs[b]) for a, b in pairs])
    assert got.tobytes() == want.tobytes()
    assert abs(got[0] - 1.0) < 1e-12 and got[len(vecs) - 1] == 0.0


def test_empty_vocabulary(orbx, tmp_path):
    """A header-only file loads (the reference accepts it) and leaves a vocabulary without words: transform returns empty vectors
    (TemplatedVocabulary.h:1137-1140)."""
    p = tmp_path / 'empty.txt'
    p.write_text('10 6 0 0\n')
    v = orbx.ORBVocabulary()
    assert v.loadFromTextFile(str(p)) and v.empty() and v.info()['nodes'] == 1
    (wi, wv), (fn, fs, fi) = v.transform(np.random.RandomState(0).randint(0, 256, (50, 32)).astype(np.uint8), 4)
    assert len(wi) == 0 and len(fn) == 0 and list(fs) == [0]


def test_loader_rejects_what_the_reference_rejects(orbx, tmp_path):
    for header in ('21 6 0 0', '10 11 0 0', '10 0 0 0', '10 6 6 0', '10 6 0 4'):
        p = tmp_path / 'bad.txt'
        p.write_text(header + '\n')
        assert orbx.ORBVocabulary().loadFromTextFile(str(p)) is False
    assert orbx.ORBVocabulary().loadFromTextFile(str(tmp_path / 'missing.txt')) is False
