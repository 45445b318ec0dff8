"""Host-side file boundary: names, dtypes and layouts the reference scripts read / write."""
import numpy as np
import pandas as pd
import pytest

from otto_multi_objective_recommender_system_b200 import io, synth


def test_part_names_and_counts_follow_the_reference_loaders():
    # covisitation/inference.py:87-111 (validation: parts 0-3, cart_order part 0), :282-308 (submission: 0-5 / 0-1)
    assert io.part_name("time_weighted", 3) == "top_15_time_weighted_3.pqt"
    assert io.part_name("cart_order", 0, None) == "top_cart_order_0.pqt"       # ranker/regular_candidate_generation.py:75-101
    assert io.n_parts_for("time_weighted", "validation") == 4 and io.n_parts_for("cart_order", "validation") == 1
    assert io.n_parts_for("cart_weighted", "submission") == 6 and io.n_parts_for("cart_order", "submission") == 2
    with pytest.raises(ValueError, match="Invalid mode"):
        io.n_parts_for("time_weighted", "train")


def test_event_frame_round_trip_parquet_and_ms_pickle(tmp_path):
    frame = synth.generate(synth.SynthSpec("train", 200, 50, seed=1))
    io.write_event_frame(frame, tmp_path / "a.parquet")
    back = io.read_event_frame(tmp_path / "a.parquet", n_aids=50)
    for c in ("session", "aid", "ts", "type"):
        assert np.array_equal(getattr(back, c).numpy(), getattr(frame, c).numpy()), c
    assert pd.read_parquet(tmp_path / "a.parquet").dtypes.astype(str).to_dict() == {
        "session": "int32", "aid": "int32", "ts": "int32", "type": "uint8"}
    # pickles carry ts in milliseconds (utilities/dataset_writer_pickle.py:57-60); consumers divide by 1000
    df = frame.to_pandas()
    df.assign(ts=df["ts"].astype(np.uint64) * 1000 + 7).to_pickle(tmp_path / "b.pkl")
    back = io.read_event_frame(tmp_path / "b.pkl", n_aids=50)
    assert np.array_equal(back.ts.numpy(), frame.ts.numpy())
    both = io.read_event_frame(tmp_path / "a.parquet", tmp_path / "b.pkl", n_aids=50)
    assert len(both) == 2 * len(frame)


def test_submission_frame_layout():
    # covisitation/inference.py:430-441: "<session>_<type>s", space-joined labels, click / cart / order adjacent
    pred = np.array([[[1, 2, -1], [3, -1, -1]], [[4, 5, 6], [7, 8, -1]], [[9, -1, -1], [10, 11, 12]]])
    f = io.submission_frame(np.array([100, 101]), pred)
    assert f["session_type"].tolist() == ["100_clicks", "100_carts", "100_orders", "101_clicks", "101_carts", "101_orders"]
    assert f["labels"].tolist() == ["1 2", "4 5 6", "9", "3", "7 8", "10 11 12"]


def test_read_popular(tmp_path):
    import json
    for e, aids in (("click", [5, 3]), ("cart", [9]), ("order", [1, 2, 4])):
        json.dump({str(a): 10 - i for i, a in enumerate(aids)}, open(tmp_path / f"train_20_most_frequent_{e}_aids.json", "w"))
    assert io.read_popular(tmp_path, "train") == {"click": [5, 3], "cart": [9], "order": [1, 2, 4]}


def test_candidate_frame_file_names(tmp_path):
    f = pd.DataFrame({"session": [1], "candidates": np.uint64([2]), "candidate_scores": np.float32([3])})
    paths = io.write_candidate_frames({"click": f, "cart": f, "order": f}, tmp_path, "submission")
    assert sorted(p.name for p in paths) == ["cart_covisitation_test.pkl", "click_covisitation_test.pkl", "order_covisitation_test.pkl"]
    back = pd.read_pickle(paths[0])
    assert back.dtypes.astype(str).to_dict() == {"session": "int64", "candidates": "uint64", "candidate_scores": "float32"}
    with pytest.raises(ValueError, match="Invalid mode"):
        io.write_candidate_frames({}, tmp_path, "nope")


def test_chunk_directory_round_trip(tmp_path):
    """utilities/split_dataset_writer_parquet.py:21-33: files of consecutive session ids, read back in chunk order."""
    import pandas as pd
    from otto_multi_objective_recommender_system_b200 import io, synth
    frame = synth.generate(synth.SynthSpec("train", 250, 40, seed=3))
    paths = io.write_event_chunks(frame, tmp_path / "train_truncated_parquet", "train_truncated", session_chunk_size=100)
    assert [p.name for p in paths] == [f"train_truncated_{i}.parquet" for i in range(3)]       # 250 // 100 + 1
    first = pd.read_parquet(paths[0])
    assert first["session"].max() < 100 and pd.read_parquet(paths[1])["session"].between(100, 199).all()
    # eleven chunks would sort 0, 1, 10, 2 ... by name: the reader orders by chunk index
    (tmp_path / "many").mkdir()
    for i in range(11):
        pd.DataFrame({"session": [i], "aid": [i], "ts": [1659304800 + i], "type": [0]}).to_parquet(tmp_path / "many" / f"c_{i}.parquet")
    assert io.read_event_frame(tmp_path / "many").to_pandas()["session"].tolist() == list(range(11))
    back = io.read_event_frame(tmp_path / "train_truncated_parquet", n_aids=40).to_pandas()
    want = frame.to_pandas().sort_values(["session", "ts"], kind="stable").reset_index(drop=True)
    assert back.astype("int64").equals(want.astype("int64"))
    import pytest
    (tmp_path / "empty").mkdir()
    with pytest.raises(FileNotFoundError):
        io.read_event_frame(tmp_path / "empty")


def test_popular_prefix_per_mode_matches_the_reference_text():
    # covisitation/inference.py:76-83 reads train_20_most_frequent_*, :271-278 reads test_20_most_frequent_*
    import pathlib
    import re
    from otto_multi_objective_recommender_system_b200 import inference
    assert inference.POPULAR_PREFIX == {"validation": "train", "submission": "test"}
    ref = pathlib.Path("/root/reference/src/covisitation/inference.py")
    if not ref.exists():
        pytest.skip("reference tree not present on this box")
    text = ref.read_text()
    val_at, sub_at = text.rindex("if args.mode == 'validation'"), text.rindex("elif args.mode == 'submission'")
    for mode, (lo, hi) in {"validation": (val_at, sub_at), "submission": (sub_at, len(text))}.items():
        prefixes = set(re.findall(r"'(\w+)_20_most_frequent_(?:click|cart|order)_aids\.json'", text[lo:hi]))
        assert prefixes == {inference.POPULAR_PREFIX[mode]}, (mode, prefixes)
