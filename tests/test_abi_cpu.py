"""The C ABI without a GPU: libotto_covisit.so builds, loads, exports every function include/otto_covisit.h
declares, the ctypes mirror of every struct has the C layout (checked against gcc), and the host-only entry
points behave (sizes, argument errors, error strings).  No kernel is launched here."""
import ctypes as C
import pathlib
import re
import subprocess

import pytest

ROOT = pathlib.Path(__file__).resolve().parents[1]
HEADER = ROOT / "include" / "otto_covisit.h"


def declared_functions():
    text = re.sub(r"/\*.*?\*/", "", HEADER.read_text(), flags=re.S)
    return sorted(set(re.findall(r"\b(otto_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_function(native_lib):
    from otto_multi_objective_recommender_system_b200 import _native as N
    names = declared_functions()
    assert len(names) >= 30
    raw = C.CDLL(str(N.LIB_PATH))
    missing = [n for n in names if not hasattr(raw, n)]
    assert not missing, f"declared in the header but not exported: {missing}"
    unbound = [n for n in names if n not in N.EXPORTED_SYMBOLS]
    assert not unbound, f"exported but not bound in _native._SIGNATURES: {unbound}"
    stale = [n for n in N.EXPORTED_SYMBOLS if n not in names]
    assert not stale, f"bound in _native but not declared in the header: {stale}"


def test_struct_layouts_match_the_c_compiler(native_lib, tmp_path):
    from otto_multi_objective_recommender_system_b200 import _native as N
    structs = ["OttoEvents", "OttoCovisitSpec", "OttoBuildSizes", "OttoBuildStats", "OttoPairSegment", "OttoTopK", "OttoSessions",
               "OttoCandidateSpec", "OttoCandidates", "OttoRecencySpec", "OttoOwnerPlan", "OttoLabels", "OttoCandidateFrame",
               "OttoInteractionFeatures"]
    lines = ['#include <stdio.h>', '#include <stddef.h>', f'#include "{HEADER}"', "int main(void) {"]
    for s in structs:
        lines.append(f'  printf("{s} %zu\\n", sizeof({s}));')
        for fname, _ in getattr(N, s)._fields_:
            lines.append(f'  printf("{s}.{fname} %zu\\n", offsetof({s}, {fname}));')
    lines += ["  return 0;", "}"]
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-std=c11", "-o", str(exe), str(src)], check=True)
    got = dict(l.split() for l in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.splitlines())
    for s in structs:
        cls = getattr(N, s)
        assert int(got[s]) == C.sizeof(cls), s
        for fname, _ in cls._fields_:
            assert int(got[f"{s}.{fname}"]) == getattr(cls, fname).offset, f"{s}.{fname}"


def test_host_only_entry_points(native_lib):
    from otto_multi_objective_recommender_system_b200 import _native as N
    from otto_multi_objective_recommender_system_b200 import covisit
    lib = native_lib
    assert lib.otto_version() >= 100
    assert isinstance(lib.otto_launch_count(), int)
    spec = covisit.CLICKS.to_c(1000)
    sizes = N.OttoBuildSizes()
    assert lib.otto_covisit_sizes(5000, 80000, C.byref(spec), C.byref(sizes)) == N.OTTO_OK
    assert sizes.tail_capacity == 80000 and sizes.max_bins > 1000 and sizes.workspace_bytes > 80000 * 12
    # a rank of a multi-GPU build sizes its bin arrays for the global frame
    import dataclasses
    big = dataclasses.replace(covisit.CLICKS, global_events=8_000_000).to_c(1000)
    sizes2 = N.OttoBuildSizes()
    assert lib.otto_covisit_sizes(5000, 80000, C.byref(big), C.byref(sizes2)) == N.OTTO_OK
    assert sizes2.max_bins > sizes.max_bins
    # argument errors come back as codes + messages, never as exceptions across the ABI
    bad = covisit.CLICKS.to_c(1000)
    bad.tail_n = 33
    assert lib.otto_covisit_sizes(5000, 80000, C.byref(bad), C.byref(sizes)) == N.OTTO_EINVAL
    assert b"tail_n" in lib.otto_last_error()
    bad = covisit.CLICKS.to_c(1000)
    bad.k = 0
    assert lib.otto_covisit_sizes(5000, 80000, C.byref(bad), C.byref(sizes)) == N.OTTO_EINVAL
    assert lib.otto_covisit_reduce_scratch_bytes(C.byref(spec), 1200, 1000) > 0
    assert lib.otto_covisit_merge_scratch_bytes(1200) > 0
    assert lib.otto_recency_scratch_bytes(458, 20) > 0
    with pytest.raises(N.OttoError) as e:
        N.check(lib.otto_covisit_sizes(5000, 1 << 31, C.byref(spec), C.byref(sizes)))
    assert e.value.code == N.OTTO_EINVAL


def test_owner_plan_argument_checks(native_lib):
    """Owner-direct entry points validate the plan on the host before any CUDA call."""
    from otto_multi_objective_recommender_system_b200 import _native as N
    from otto_multi_objective_recommender_system_b200 import covisit
    lib = native_lib
    spec = covisit.CLICKS.to_c(1000)
    ev = N.OttoEvents(10, 100, None, None, None, None)

    def plan(n, rank, cuts):
        p = N.OttoOwnerPlan()
        p.n_owners, p.rank = n, rank
        for i in range(N.MAX_OWNERS + 1):
            p.aid_cuts[i] = cuts[min(i, len(cuts) - 1)]
        return p
    for bad, what in ((plan(0, 0, [0, 1000]), b"n_owners"), (plan(9, 0, [0, 1000]), b"n_owners"), (plan(2, 2, [0, 500, 1000]), b"rank"),
                      (plan(2, 0, [0, 500, 900]), b"aid_cuts"), (plan(2, 0, [0, 1200, 1000]), b"non-decreasing")):
        assert lib.otto_covisit_count_finish_owned(C.byref(ev), C.byref(spec), None, 0, C.byref(bad), None, None, None) == N.OTTO_EINVAL
        assert what in lib.otto_last_error(), lib.otto_last_error()
    ok = plan(2, 1, [0, 500, 1000])
    assert lib.otto_covisit_count_finish_owned(C.byref(ev), C.byref(spec), None, 0, C.byref(ok), None, None, None) == N.OTTO_EINVAL
    assert b"row_before" in lib.otto_last_error()
    # the scatter needs every owner's buffer
    assert lib.otto_covisit_scatter_owned(C.byref(ev), C.byref(spec), None, 0, C.byref(ok), None) in (N.OTTO_EINVAL, N.OTTO_ENOSPC)
    assert lib.otto_covisit_count_finish_owned(None, C.byref(spec), None, 0, C.byref(ok), None, None, None) == N.OTTO_EINVAL


def test_product_path_has_no_cpu_fallback():
    """CPU tensors are rejected before any work: there is no host implementation to fall back to."""
    import torch
    from otto_multi_objective_recommender_system_b200 import covisit, synth
    frame = synth.generate(synth.SynthSpec("train", 50, 20, seed=1))
    csr = covisit.EventCSR(torch.arange(3, dtype=torch.int32), torch.tensor([0, 1, 2, 3], dtype=torch.int32), frame.aid[:3],
                           frame.ts[:3], frame.type[:3], 20, "desc")
    with pytest.raises(RuntimeError, match="CUDA device"):
        covisit.CovisitBuilder(csr, covisit.CLICKS)
