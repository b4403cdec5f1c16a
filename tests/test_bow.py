"""Frame::ComputeBoW (reference src/Frame.cc:1692-1699) = DBoW2::TemplatedVocabulary::transform(features, BowVector, FeatureVector, levelsup),
pinned BY EXECUTION of the reference's own DBoW2 sources (oracle/_ref/ref_bow: Thirdparty/DBoW2 compiled unmodified; ORBvoc.bin is absent
from the reference tree, so the vocabulary is built by DBoW2's own create() from synthetic ORB descriptors and handed over as arrays):

  CPU: oracle (oracle/bow_oracle.cpp) == executed reference (live, or the committed fixture tests/golden/bow_ref.npz)
  GPU: hvo_bow_transform == oracle == reference: word ids, the L1-normalised BowVector values (doubles compared by their bytes), the
       FeatureVector node -> feature lists."""
import os

import numpy as np
import pytest

import oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, 'tests', 'golden', 'bow_ref.npz')
K, L, LEVELSUP = 10, 4, 2


def _descriptors(synth):
    ex = oracle.OrbOracle()
    train = [ex.extract(synth.frame('S1', 10 + i)[0])[1] for i in range(24)]
    frames = [ex.extract(synth.frame('S1', i)[0])[1] for i in range(3)] + [ex.extract(synth.frame('S2', 0)[0])[1]]
    frames.append(frames[0][:1])          # one feature
    frames.append(frames[1][:0])          # no features
    return train, frames


def make_golden(synth):
    train, frames = _descriptors(synth)
    voc, res = oracle.ref_bow(train, frames, K, L, 1, LEVELSUP)
    out = {f'voc_{k}': np.asarray(v) for k, v in voc.items()}
    for i, ((w, v), fv) in enumerate(res):
        out[f'f{i}_words'] = w; out[f'f{i}_values'] = v
        out[f'f{i}_fv_nodes'] = np.asarray(sorted(fv), np.int32)
        out[f'f{i}_fv_feats'] = np.asarray([x for k in sorted(fv) for x in fv[k]], np.int32)
        out[f'f{i}_fv_counts'] = np.asarray([len(fv[k]) for k in sorted(fv)], np.int32)
    np.savez_compressed(GOLDEN, **out)
    return out


@pytest.fixture(scope='module')
def ref(synth):
    """(voc, [((words, values), fv)], frames): the executed reference, live when oracle/_ref/ref_bow exists (checked against the fixture)."""
    train, frames = _descriptors(synth)
    g = np.load(GOLDEN) if os.path.exists(GOLDEN) else None
    live = oracle.ref_bow(train, frames, K, L, 1, LEVELSUP)
    if live is None and g is None:
        pytest.skip('neither oracle/_ref/ref_bow nor tests/golden/bow_ref.npz is available')
    if g is not None:
        voc = {k[4:]: g[k] for k in g.files if k.startswith('voc_')}
        voc['L'] = int(voc['L'])
        res = []
        for i in range(len(frames)):
            ends = np.cumsum(g[f'f{i}_fv_counts'])
            fv = {int(n): list(map(int, g[f'f{i}_fv_feats'][e - c:e])) for n, e, c in zip(g[f'f{i}_fv_nodes'], ends, g[f'f{i}_fv_counts'])}
            res.append(((g[f'f{i}_words'], g[f'f{i}_values']), fv))
        if live is not None:     # the fixture is what the reference produces here
            lvoc, lres = live
            for k in ('child_start', 'child_ids', 'node_desc', 'node_weight', 'node_word'):
                assert np.asarray(lvoc[k]).tobytes() == np.asarray(voc[k]).tobytes(), f'fixture vocabulary differs from the live reference: {k}'
            for (a, fa), (b, fb) in zip(lres, res):
                assert np.array_equal(a[0], b[0]) and a[1].tobytes() == np.asarray(b[1]).tobytes() and fa == fb
        return voc, res, frames
    return live[0], live[1], frames


def _check(got, want):
    (gw, gv), gfv = got[0], got[1]
    (ww, wv), wfv = want
    assert np.array_equal(gw, ww)
    assert np.asarray(gv, np.float64).tobytes() == np.asarray(wv, np.float64).tobytes()
    assert gfv == wfv


def test_oracle_bow_equals_reference(ref):
    voc, res, frames = ref
    assert len(voc['node_weight']) > 2000 and len(res[0][0][0]) > 300
    assert len(res[0][1]) > 20 and max(len(v) for v in res[0][1].values()) > 3   # FeatureVector groups several features per node
    assert abs(res[0][0][1].sum() - 1.0) < 1e-9
    for d, want in zip(frames, res):
        _check(oracle.bow_transform(voc, d, LEVELSUP), want)


@pytest.mark.gpu
def test_gpu_bow_equals_oracle_and_reference(hvo, ref):
    voc, res, frames = ref
    v = hvo.ORBVocabulary(voc)
    got = v.transform_batch(frames, LEVELSUP)
    for d, g, want in zip(frames, got, res):
        _check(g, want)
        o = oracle.bow_transform(voc, d, LEVELSUP)
        assert np.array_equal(g[2], o[2]) and np.array_equal(g[3], o[3])
    # other levelsup values (nid at other levels, incl. the root for levelsup >= L), one frame at a time
    for lu in (0, 1, 3, 4, 6):
        g = v.transform(frames[0], lu)
        o = oracle.bow_transform(voc, frames[0], lu)
        _check(g, (o[0], o[1]))
        assert np.array_equal(g[3], o[3])
    v.close()


@pytest.mark.gpu
def test_dropin_bow_transformer_equals_executed_reference(synth, ref):
    """oracle/_ref/shim_bow = the executed reference's own driver (oracle/ref_bow_main.cpp) and DBoW2's own vocabulary object, built by its
    create(), with transform(features, BowVector, FeatureVector, levelsup) routed through the drop-in hvo_shim::BowTransformerT
    (shim/ORBVocabularyGPU.h -> C ABI -> CUDA): it must write what the reference's transform wrote.  The binary compiles DBoW2's sources, so it
    is built where the reference tree is mounted (oracle/Makefile) and travels with oracle/_ref."""
    if oracle.ref_bin('shim_bow') is None:
        pytest.skip('oracle/_ref/shim_bow is not built (reference tree absent at build time)')
    voc, res, frames = ref
    train, frames2 = _descriptors(synth)
    svoc, sres = oracle.ref_bow(train, frames2, K, L, 1, LEVELSUP, exe_name='shim_bow')
    for k in ('child_start', 'child_ids', 'node_desc', 'node_weight', 'node_word'):
        assert np.asarray(svoc[k]).tobytes() == np.asarray(voc[k]).tobytes()
    assert len(sres) == len(res)
    for got, want in zip(sres, res):
        _check(got, want)
