"""CPU-side checks (no GPU): the C-ABI library loads and exports every symbol include/msa_b200.h
declares, capacity queries work without a device, host-side sharding / row packing logic, the
reference-compatible state_dict layout, and the loud failure of compute calls without CUDA."""
import os
import re

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _host_table(audio_row, logits, first_id):
    """The [n, 40] result-table layout (msa_b200.pipeline / msa_pack_rows) built on the host for the CPU tests."""
    from msa_b200.pipeline import ROW_WORDS
    n = audio_row.shape[0]
    rows = torch.empty(n, ROW_WORDS, dtype=torch.float32)
    rows[:, 0:31], rows[:, 31:38] = audio_row, logits
    ints = rows.view(torch.int32)
    ints[:, 38] = logits.argmax(1).to(torch.int32)
    ints[:, 39] = torch.arange(first_id, first_id + n, dtype=torch.int32)
    return rows


@pytest.fixture(scope="module")
def built():
    import __graft_entry__ as g
    g.build()
    import msa_b200
    return msa_b200


def test_header_symbols_are_exported(built):
    from msa_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "msa_b200.h")).read()
    declared = set(re.findall(r"\b(msa_[a-z0-9_]+)\s*\(", hdr))
    l = _lib.lib()
    for name in declared:
        assert hasattr(l, name), f"{name} declared in include/msa_b200.h but not exported"
    assert declared == set(_lib.exported_symbols())


def test_capacity_queries_without_device(built):
    from msa_b200 import _lib
    l = _lib.lib()
    assert l.msa_version() >= 100
    assert l.msa_features_cluster_size(80000) == 1                        # a 5 s segment fits one CTA
    assert l.msa_features_cluster_size(8000) == 1
    assert l.msa_features_cluster_size(160000) == 2                       # smallest cluster that still allows two CTAs per SM
    assert l.msa_features_cluster_size(960000) in (4, 8)                  # one minute: one CTA per SM, split over a cluster
    assert l.msa_features_cluster_size(2_000_000) == 8
    assert l.msa_features_cluster_size(4_000_000) == 0                    # unsupported: too long for one cluster
    assert 0 < l.msa_features_smem_bytes(80000, 4) <= 232448
    assert l.msa_fusion_num_tensors() == 42
    assert l.msa_strerror(0) == b"ok" and b"too long" in l.msa_strerror(-2)
    assert l.msa_fusion_workspace_bytes(1024) > 0 and l.msa_fusion_packed_bytes() > 4 * 5_600_000
    # the activation workspace grows with the batch, also across the 3072-row switch from 128-row tiles (one CTA) to
    # 256-row tiles (CTA pairs): a workspace sized for the largest batch serves every smaller one
    sizes = [l.msa_fusion_workspace_bytes(b) for b in (1, 128, 129, 3072, 3073, 3200, 4096, 5000, 65536)]
    assert all(a <= b for a, b in zip(sizes, sizes[1:])), sizes
    assert l.msa_fusion_workspace_bytes(3073) == l.msa_fusion_workspace_bytes(3328)      # both pad to 13 tiles of 256 rows


def test_state_dict_matches_reference_layout(built):
    """45 tensors, 5,604,508 parameters, reference key names."""
    from msa_b200 import _lib
    from oracle import synth
    m = built.AdvancedFusionModel(device="cpu")
    sd = m.state_dict()
    assert len(sd) == 45 and sum(v.numel() for v in sd.values()) == 5_604_508
    assert set(sd.keys()) == set(synth.fusion_state(0).keys())
    l = _lib.lib()
    names = [l.msa_fusion_tensor_name(i).decode() for i in range(l.msa_fusion_num_tensors())]
    assert set(names) == set(sd.keys()) - {"audio_weight", "text_weight", "face_weight"}
    for i, n in enumerate(names):
        assert l.msa_fusion_tensor_numel(i) == sd[n].numel()
    w = m.get_weights()                                                    # 0.3 / 0.3 / 0.4 -> softmax
    assert abs(w["audio"] - 0.3220) < 1e-3 and abs(w["face"] - 0.3559) < 1e-3
    # single-modality pass-through and the always-failing pairs need no device
    a = torch.randn(2, 31)
    assert m(None, a, None)["audio"] is a
    assert list(m(None, a, torch.randn(2, 783)).keys()) == ["audio"]


def test_no_cpu_fallback(built):
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from msa_b200 import _lib
    with pytest.raises(_lib.MsaError):
        built.AudioAnalyzer(device="cuda")
    m = built.AdvancedFusionModel(device="cpu")
    with pytest.raises(_lib.MsaError):
        m(torch.randn(2, 27), torch.randn(2, 31), torch.randn(2, 783))


def test_shard_ranges_and_row_packing(built):
    from msa_b200.pipeline import pack_rows, shard_range, unpack_rows
    for n, w in ((720, 8), (1024, 3), (5, 8), (0, 2)):
        r = [shard_range(n, w, k) for k in range(w)]
        assert r[0][0] == 0 and r[-1][1] == n and all(r[i][1] == r[i + 1][0] for i in range(w - 1))
        sizes = [e - b for b, e in r]
        assert max(sizes) - min(sizes) <= 1
    lg = torch.randn(5, 7)
    u = unpack_rows(_host_table(torch.randn(5, 31), lg, 1 << 20))
    assert u["argmax"].tolist() == lg.argmax(1).tolist() and u["segment_id"].tolist() == list(range(1 << 20, (1 << 20) + 5))
    with pytest.raises(Exception):
        pack_rows(torch.randn(5, 31), lg, lg.argmax(1), 0)                  # device only: no CPU path in the product


def _gather_worker(rank, world, port, n_total, out_dir):
    import os
    import numpy as np
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import msa_b200  # noqa: F401
    from msa_b200.pipeline import ROW_WORDS, gather_rows, gather_rows_async, shard_range, unpack_rows
    b, e = shard_range(n_total, world, rank)
    g = torch.Generator().manual_seed(1000)                      # every rank draws the same full table ...
    full_audio, full_logits = torch.randn(n_total, 31, generator=g), torch.randn(n_total, 7, generator=g)
    rows = _host_table(full_audio[b:e], full_logits[b:e], b)                   # ... and owns one shard of it
    table = gather_rows(rows, n_total, world, rank)
    u = unpack_rows(table)
    assert table.shape == (n_total, ROW_WORDS)
    assert torch.equal(u["audio_row"], full_audio) and torch.equal(u["logits"], full_logits)
    assert u["segment_id"].tolist() == list(range(n_total))
    assert torch.equal(u["argmax"].long(), full_logits.argmax(1))
    pend = gather_rows_async(rows, n_total, world, rank)                      # the overlapped form gives the same table
    assert torch.equal(pend.wait().view(torch.int32), table.view(torch.int32))
    np.save(os.path.join(out_dir, f"ok{rank}.npy"), np.array([table.shape[0]]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_total", [720, 7, 1])
def test_sharded_gather_world_size_2_gloo(built, tmp_path, n_total):
    """contiguous shards, ONE all_gather of the [S/N, 40] result tables (ragged shards
    padded and trimmed).  Two CPU processes over gloo stand in for two GPUs over NCCL."""
    import socket
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_gather_worker, args=(2, port, n_total, str(tmp_path)), nprocs=2, join=True)
    assert os.path.exists(os.path.join(tmp_path, "ok0.npy")) and os.path.exists(os.path.join(tmp_path, "ok1.npy"))
