import sys, time
sys.path.insert(0, "/root/repo")
import torch
import __graft_entry__ as g
g.build()
from otto_multi_objective_recommender_system_b200 import candidates, covisit, synth
dev = torch.device("cuda:0")
frame = synth.generate(synth.SynthSpec.scaled("train", 1.0), device=dev)
csr = covisit.ingest(frame, "desc", device=dev)
A = csr.n_aids
tables = {stem: covisit.build_topk(csr, spec)[0] for stem, spec in covisit.VARIANTS.items()}
del frame, csr
test = synth.generate(synth.SynthSpec.scaled("test", 1.0), device=dev)
sess = covisit.ingest(test, "asc", device=dev)
gen = candidates.CandidateGenerator(tables, candidates.reference_spec(tables.keys(), 20), A)
mlen = candidates.max_session_len(sess)
popular = {t: list(range(20)) for t in ("click", "cart", "order")}
def run():
    cand = gen(sess, mlen)
    pred, long_s = candidates.assemble_predictions(sess, cand, popular, 20)
    candidates.recency_long_predictions(sess, tables, pred, long_s, 20)
run(); torch.cuda.synchronize()
torch.cuda.nvtx.range_push("step"); run(); torch.cuda.synchronize(); torch.cuda.nvtx.range_pop()
