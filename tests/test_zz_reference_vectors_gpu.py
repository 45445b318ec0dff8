"""CUDA path vs vectors produced by the reference's OWN loop bodies (tests/golden/reference_candidates.json, made by
tests/golden/make_reference_vectors.py).  The file name sorts last on purpose: the checks before it compare against
the oracle on seeded inputs; this one closes the chain reference code -> golden vectors -> CUDA path."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def mods(native_lib):
    from otto_multi_objective_recommender_system_b200 import candidates, covisit, synth
    return covisit, candidates, synth


@pytest.mark.parametrize("vectors", ["reference_candidates.json", "reference_candidates_wide.json"])
def test_cuda_path_matches_vectors_from_the_reference_loop_bodies(mods, vectors):
    """tests/golden/reference_candidates.json holds outputs of the reference's OWN loop bodies (executed as text by
    tests/golden/make_reference_vectors.py): ranker form, regular form, recency-weighted generator and the standalone
    model with its long-session branch, all seven stems - on 60 short sessions with rows of <= 6 neighbours, and on the
    wide set (K = 15 rows, sessions of up to 458 events, most_common(100) truncating).  The CUDA path must reproduce them."""
    import parity_helpers as H
    cv, cand_mod, synth = mods
    g = H.reference_vectors(vectors)
    df, otables, labels, popular = H.reference_vector_inputs(g)
    n_aids = g["n_aids"]
    tables = {}
    for stem, rows in otables.items():
        r = H.table_rows(rows)
        tables[stem] = cv.TopKTable.from_rows(torch.tensor(r["aid_x"].to_numpy(), device="cuda:0"),
                                              torch.tensor(r["aid_y"].to_numpy(), device="cuda:0"),
                                              torch.tensor(r["wgt"].to_numpy(), dtype=torch.float32, device="cuda:0"), n_aids, g.get("table_k", 6))
    sess = cv.ingest(synth.EventFrame.from_pandas(df, n_aids), "asc", device="cuda:0")
    ranker = cand_mod.generate_candidates(sess, tables, cand_mod.reference_spec(tables.keys(), 100)).to_frames(labels)
    H.check_candidate_frames(g, "ranker", ranker)
    H.check_candidate_frames(g, "regular", cand_mod.regular_candidates(sess, tables, 100, labels=labels))
    rw = cand_mod.recency_weighted_candidates(sess, labels=labels, keep_f64=True)
    H.check_candidate_frames(g, "recency", rw)
    H.check_candidate_frames(g, "recency", rw, score_column="candidate_scores_f64")
    cand = cand_mod.generate_candidates(sess, tables, cand_mod.reference_spec(tables.keys(), 20))
    pred, long_session = cand_mod.assemble_predictions(sess, cand, popular, 20)
    cand_mod.recency_long_predictions(sess, tables, pred, long_session, 20)
    assert int(long_session.sum()) == sum(len(set(s["aid"])) >= 20 for s in g["sessions"])
    p = pred.cpu().numpy()
    preds = {e: [[int(a) for a in p[ti, i] if a >= 0] for i in range(sess.n_sessions)] for ti, e in enumerate(cand.targets)}
    H.check_standalone_predictions(g, preds)
