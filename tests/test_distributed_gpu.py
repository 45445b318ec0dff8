"""Multi-GPU build == single-GPU build, byte for byte (needs >= 2 GPUs: run with `gpurun --gpus 2`)."""
import os
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, variant, split_ub, n_sessions, n_aids, transport):
    import torch.distributed as dist
    # owner_direct: NVLink stores from the scatter kernel; staged: runs combined in the sender's staging buffer, owners pull
    # their buckets; peer_read: owners read the senders' slabs; nccl: all-to-all
    use_peer = transport != "nccl"
    os.environ["OTTO_OWNER_DIRECT"] = "1" if transport in ("owner_direct", "staged") else "0"
    os.environ["OTTO_STAGED"] = "1" if transport == "staged" else "0"
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    from dataclasses import replace
    from otto_multi_objective_recommender_system_b200 import covisit, distributed, synth
    spec = replace(getattr(covisit, variant), split_ub=split_ub)
    frame = synth.generate(synth.SynthSpec("train", n_sessions, n_aids, seed=31), device=dev)
    csr = covisit.ingest(frame, "desc", device=dev)
    S = csr.n_sessions
    shard = csr.slice_sessions(rank * S // world, (rank + 1) * S // world)
    peer = distributed.PeerRecords(dev) if use_peer else None     # NVLink peer memory vs NCCL all-to-all
    backend = distributed.GpuRankBackend(shard, spec, exact=True, peer=peer)
    assert backend.owner_direct == (transport in ("owner_direct", "staged")) and backend.staged == (transport == "staged")
    table, (lo, hi), stats, plan = distributed.build_topk_distributed(backend)
    distributed.gather_table(table, plan)
    single, sstats = covisit.build_topk(csr, spec, exact=True)
    for name in ("aid_y", "wgt", "len"):
        assert torch.equal(getattr(table, name), getattr(single, name)), (rank, name)
    # exact integer side outputs on the rows this rank owns
    assert torch.equal(table.cnt[lo:hi], single.cnt[lo:hi]) and torch.equal(table.tsum[lo:hi], single.tsum[lo:hi])
    total = torch.tensor([stats["pair_checksum"], stats["distinct"]], device=dev)
    dist.all_reduce(total)
    assert int(total[0]) == sstats["pair_checksum"] and int(total[1]) == sstats["distinct"]
    if peer is not None:
        peer.close()
    dist.destroy_process_group()


@pytest.mark.parametrize("transport", ["owner_direct", "staged", "peer_read", "nccl"])
@pytest.mark.parametrize("variant,split_ub", [("CLICKS", 0), ("CARTS_ORDERS", 64), ("BUY2BUY", 0)])
def test_multi_gpu_equals_single_gpu(native_lib, variant, split_ub, transport):
    world = min(torch.cuda.device_count(), 8)      # every GPU of the box: the owner-direct scatter must hold at 8 ranks too
    if world < 2:
        pytest.skip("needs at least 2 GPUs")
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    torch.multiprocessing.spawn(_worker, args=(world, port, variant, split_ub, 20000, 2500, transport), nprocs=world, join=True)


@pytest.mark.parametrize("transport", ["staged", "owner_direct"])
def test_eight_rank_owner_direct_full_width(native_lib, transport):
    """The default transport (staged) and the direct scatter at the full width of a box (8 ranks, or whatever the box has
    beyond 4): a larger frame, so that every owner holds hot rows and receives records from every peer."""
    world = min(torch.cuda.device_count(), 8)
    if world <= 4:
        pytest.skip("needs more than 4 GPUs")
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    torch.multiprocessing.spawn(_worker, args=(world, port, "CLICKS", 256, 60000, 4000, transport), nprocs=world, join=True)
