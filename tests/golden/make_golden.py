"""Regenerates the golden fixtures under tests/golden/ (run from the repo root, in the build container).

  session_747.json   the one real OTTO session printed in the reference's EDA notebook
                     (notebook/otto-multi-objective-recommender-system-eda.ipynb, cell 37 output):
                     29 events (aid, ts, type); ts floored from ms to s as the consumers do
                     (ranker/aid_feature_engineering.py:36).  Read from /root/reference when present.
  oracle_small.json  a 300-session synthetic frame and the three top-K tables the pandas oracle
                     (oracle/covisit_oracle.py) yields for it - regression vectors for the oracle itself.
  popular.json       the train top-20 click / cart / order aids shipped with the reference
                     (data/aid_frequencies/train_20_most_frequent_{click,cart,order}_aids.json), used as
                     the popular-fill lists of covisitation/inference.py:76-83.
"""
import json
import pathlib
import sys
from datetime import datetime, timezone

ROOT = pathlib.Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
REF = pathlib.Path("/root/reference")
OUT = pathlib.Path(__file__).resolve().parent


def session_747():
    nb = json.load(open(REF / "notebook" / "otto-multi-objective-recommender-system-eda.ipynb"))
    cell = next(c for c in nb["cells"] if c["cell_type"] == "code" and "df_session747 = " in "".join(c["source"]))
    text = "".join(cell["outputs"][0]["data"]["text/plain"])
    rows = []
    for line in text.splitlines():
        parts = line.split()
        if len(parts) >= 6 and parts[1] == "747":
            aid, date, clock, typ = int(parts[2]), parts[3], parts[4], int(parts[5])
            dt = datetime.strptime(f"{date} {clock}", "%Y-%m-%d %H:%M:%S.%f").replace(tzinfo=timezone.utc)
            rows.append({"aid": aid, "ts": int(dt.timestamp()), "type": typ})
        if len(rows) == 29:
            break
    assert len(rows) == 29
    json.dump({"session": 747, "events": rows}, open(OUT / "session_747.json", "w"), indent=1)


def popular():
    out = {}
    for name in ("click", "cart", "order"):
        d = json.load(open(REF / "data" / "aid_frequencies" / f"train_20_most_frequent_{name}_aids.json"))
        out[name] = [int(a) for a in d.keys()]
    json.dump(out, open(OUT / "popular.json", "w"))


def oracle_small():
    from otto_multi_objective_recommender_system_b200 import synth
    from oracle import covisit_oracle as co
    frame = synth.generate(synth.SynthSpec("train", 300, 120, seed=5))
    df = frame.to_pandas()
    out = {"frame": {c: df[c].tolist() for c in df.columns}, "tables": {}}
    for name, spec in (("clicks", co.CLICKS), ("carts_orders", co.CARTS_ORDERS), ("buy2buy", co.BUY2BUY)):
        t = co.build(df, spec)
        out["tables"][name] = {"aid_x": t["aid_x"].tolist(), "aid_y": t["aid_y"].tolist(),
                               "wgt": [float(w) for w in t["wgt"]]}
    json.dump(out, open(OUT / "oracle_small.json", "w"))


if __name__ == "__main__":
    session_747()
    popular()
    oracle_small()
