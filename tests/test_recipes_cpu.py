"""Recipe files for the matrix stems (host logic only)."""
import json
import pathlib

import pytest

ROOT = pathlib.Path(__file__).resolve().parents[1]


def test_example_recipes_parse_and_cover_the_unpinned_stems():
    from otto_multi_objective_recommender_system_b200 import candidates, covisit, _native as N
    recipes = covisit.load_stem_recipes(ROOT / "configs" / "unpinned_stems.example.json")
    assert set(recipes) | set(covisit.VARIANTS) == set(candidates.STEMS)      # all seven files of covisitation/inference.py:87-111
    assert recipes["click_weighted"].weight_mode == N.WEIGHT_TYPE and recipes["click_weighted"].type_weight == (6, 3, 1)
    assert recipes["click_cart"].event_types == (0, 1) and recipes["click_cart"].window_s == 14 * 86400
    c = recipes["click_order"].to_c(100)
    assert c.event_type_mask == 0b101 and c.x_type_mask == 0b111 and c.k == 15


def test_from_dict_rejects_bad_recipes(tmp_path):
    from otto_multi_objective_recommender_system_b200 import covisit
    with pytest.raises(ValueError):
        covisit.CovisitSpec.from_dict({"weight": "tfidf"})
    with pytest.raises(ValueError):
        covisit.CovisitSpec.from_dict({"weight": "type", "type_weight": [1, 2]})
    with pytest.raises(ValueError):
        covisit.CovisitSpec.from_dict({"weight": "unit", "event_types": [1, 3]})
    with pytest.raises(ValueError):
        covisit.CovisitSpec.from_dict({"weight": "unit", "top": 15})
    bad = tmp_path / "r.json"
    bad.write_text(json.dumps({"buy_weighted": {"weight": "unit"}}))
    with pytest.raises(ValueError):
        covisit.load_stem_recipes(bad)
    assert covisit.CovisitSpec.from_dict({}) == covisit.CLICKS
