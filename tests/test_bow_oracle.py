"""CPU tests of the bag-of-words oracle: the stand-alone restatement (oracle/bow_oracle.cc) against the reference's own
DBoW2 text (oracle/_ref, Makefile rule bow_gen.cc) and against the committed outputs of that text (tests/golden/bow.npz)."""
import os

import numpy as np
import pytest

import bow_cases
from orb_slam2_refactored_b200 import synth

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def _port_result(oracle_port, voc, feats, levelsup):
    v = oracle_port.vocabulary(arrays=voc)
    wi, wv, fv = v.transform(feats, levelsup)
    return bow_cases.flatten(((wi, wv), fv)), v


@pytest.mark.parametrize('name', [c[0] for c in bow_cases.CASES])
def test_port_equals_reference_text(oracle_port, oracle_ref, name, tmp_path):
    voc, feats, levelsup = bow_cases.make(name)
    path = str(tmp_path / 'voc.txt')
    synth.write_vocabulary_text(voc, path)
    ref = oracle_ref.vocabulary(path=path)
    wi, wv, fv = ref.transform(feats, levelsup)
    a = bow_cases.flatten(((wi, wv), fv))
    b, pv = _port_result(oracle_port, voc, feats, levelsup)
    assert bow_cases.same(a, b)
    # the port's own text loader (strtok/atoi, fractional weights truncated) gives the same tree as the arrays
    wi2, wv2, fv2 = oracle_port.vocabulary(path=path).transform(feats, levelsup)
    assert bow_cases.same(a, bow_cases.flatten(((wi2, wv2), fv2)))
    if name == 'k2L10_all_stopped':
        assert len(a['word_ids']) == 0 and len(a['fv_nodes']) == 0
    else:
        assert len(a['word_ids']) > 50 and len(a['fv_nodes']) >= 1
    # score of this frame against perturbed copies of itself and an unrelated frame
    if voc['scoring'] == 0:
        other = synth.vocabulary_features(999, voc, len(feats))
        wo, vo, _ = ref.transform(other, levelsup)
        for ia, va, ib, vb in ((wi, wv, wi, wv), (wi, wv, wo, vo), (wo, vo, wi[::2], wv[::2]), (wi[:0], wv[:0], wo, vo)):
            assert ref.score((ia, va), (ib, vb)) == pv.score((ia, va), (ib, vb))


def test_loader_rejects_what_the_reference_rejects(oracle_port, oracle_ref, tmp_path):
    for header in ('21 6 0 0', '10 11 0 0', '10 0 0 0', '10 6 6 0', '10 6 0 4'):
        p = tmp_path / 'bad.txt'
        p.write_text(header + '\n')
        for o in (oracle_ref, oracle_port):
            with pytest.raises(ValueError):
                o.vocabulary(path=str(p))


@pytest.mark.parametrize('name', bow_cases.SMALL)
def test_port_equals_golden(oracle_port, name):
    g = np.load(os.path.join(G, 'bow.npz'))
    voc, feats, levelsup = bow_cases.make(name)
    b, _ = _port_result(oracle_port, voc, feats, levelsup)
    for k in ('word_ids', 'word_vals', 'fv_nodes', 'fv_start', 'fv_items'):
        assert g[f'{name}_{k}'].tobytes() == b[k].tobytes(), k
    assert g[f'{name}_features'].tobytes() == feats.tobytes()     # the generator itself is pinned too
