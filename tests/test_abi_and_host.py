"""CPU: the C-ABI library loads and exports every symbol include/vince_b200.h declares (no compute calls), argument
validation fails loudly, and the host-side logic (plans, ring arithmetic, API shapes) behaves like the reference."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

import vince_oracle as vo
from conftest import ROOT, make_args


def declared_functions():
    src = open(os.path.join(ROOT, "include", "vince_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(vince_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from vince_b200 import _lib
    dll = _lib.lib()
    names = declared_functions()
    assert len(names) >= 20
    for n in names:
        assert hasattr(dll, n), "missing export: " + n
    assert set(names) == set(_lib.SIGNATURES), (set(names) ^ set(_lib.SIGNATURES))
    assert dll.vince_abi_version() == 4
    assert int(dll.vince_infonce_workspace_bytes(256, 128)) > 2 * 256 * 128 * 4


def test_struct_layouts_match_the_header():
    """sizes that nvcc's static_asserts also pin on the C side"""
    from vince_b200 import _lib
    assert ctypes.sizeof(_lib.BnSide) == 16
    assert ctypes.sizeof(_lib.ConvDesc) % 8 == 0 and ctypes.sizeof(_lib.ConvDesc) >= 200
    assert _lib.ConvDesc.stats.offset % 8 == 0 and _lib.ConvDesc.bn_counter.offset % 8 == 0


def test_argument_validation_fails_loudly_without_touching_the_gpu():
    from vince_b200 import _lib
    dll = _lib.lib()
    d = _lib.ConvDesc()
    assert dll.vince_conv_fwd(ctypes.byref(d), None) == -1
    assert "empty problem" in _lib.last_error()
    assert dll.vince_conv_fwd(None, None) == -1
    n = _lib.InfoNceDesc()
    assert dll.vince_infonce_fwd(ctypes.byref(n), None) == -1
    assert dll.vince_stem_pack(None, None, None, None, 1, 8, 8, None) == -1
    with pytest.raises(RuntimeError):
        _lib.check(-1, "x")


def test_ops_reject_cpu_tensors_and_missing_library():
    from vince_b200 import ops
    with pytest.raises(RuntimeError, match="CUDA tensors only"):
        ops.l2_normalize(torch.zeros(4, 8), torch.zeros(4, 8))
    with pytest.raises(RuntimeError, match="CUDA"):
        ops.round_tf32(torch.zeros(4), torch.zeros(4))
    code = ("import os,sys; sys.path.insert(0, %r); os.environ['VINCE_B200_LIB']='/nonexistent/lib.so';"
            "from vince_b200 import _lib\ntry:\n    _lib.lib()\nexcept RuntimeError as e:\n    print('RAISED', 'no CPU' in str(e))" % ROOT)
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=120)
    assert "RAISED True" in out.stdout, out.stdout + out.stderr


def test_new_entry_points_validate_on_the_host():
    """uint8 stem packing, the InfoNCE backward and the prefetcher refuse CPU tensors / devices before any launch."""
    from vince_b200 import _lib, ops
    from vince_b200.prefetch import BatchPrefetcher
    x8 = torch.zeros((2, 8, 8, 3), dtype=torch.uint8)
    hi = torch.zeros((2, 7, 7, 16), dtype=torch.float16)
    with pytest.raises(RuntimeError, match="CUDA"):
        ops.build_stem_pack_u8(x8, None, hi, hi.clone())
    with pytest.raises(ValueError):
        ops.build_stem_pack_u8(torch.zeros((2, 8, 8, 4), dtype=torch.uint8), None, hi, hi.clone())
    with pytest.raises(RuntimeError, match="CUDA"):
        BatchPrefetcher("cpu")
    lib = _lib.lib()
    assert lib.vince_infonce_bwd_workspace_bytes(256, 128) >= 2 * 148 * 128 * 128 * 4
    d = _lib.InfoNceDesc()
    d.B, d.Bk, d.K, d.D, d.num_frames, d.temperature = 8, 8, 0, 48, 2, 0.07          # D not a multiple of 32
    assert lib.vince_infonce_bwd(ctypes.byref(d), 1.0, 0, 0, None, None) < 0
    assert b"embedding size" in lib.vince_last_error()


def test_model_refuses_cpu_inputs_and_cpu_parameters():
    import vince_b200
    args = make_args(batch_size=4, num_frames=2, queue_size=16, embedding_size=32, device="cuda:0")
    m = vince_b200.VinceModel(args)
    with pytest.raises(RuntimeError, match="GPU"):
        m.get_embeddings({"data": torch.zeros(4, 3, 32, 32)})
    with pytest.raises(RuntimeError):
        vince_b200.StorageQueue(16, 32, device="cpu")


@pytest.mark.parametrize("backbone,n_convs,n_params", [("ResNet18", 20, 12017832), ("ResNet50", 53, 30015656)])
def test_parameter_tree_matches_the_reference(backbone, n_convs, n_params):
    import vince_b200
    args = make_args(backbone=backbone, batch_size=8, num_frames=2, queue_size=64, embedding_size=128, device="cpu")
    m = vince_b200.VinceModel(args)
    sd = vo.make_state_dict(backbone, 128)                 # key names / shapes of the reference's state_dict
    assert list(m.state_dict().keys()) == list(sd.keys())
    assert all(m.state_dict()[k].shape == sd[k].shape for k in sd)
    assert sum(p.numel() for p in m.vince_parameters()) == n_params      # a10
    runner = m.feature_extractor.module.runner
    assert len(runner.bank.specs) == n_convs
    assert m.output_channels == (512 if backbone == "ResNet18" else 2048)
    assert m.to("cpu") is None                             # reference quirk: .to() returns None (vince_model.py:92-94)
    # deep copy (VinceQueueModel) re-binds the runner to the copied parameters and freezes them
    qm = vince_b200.VinceQueueModel(args, m)
    r2 = qm.queue_network.feature_extractor.module.runner
    assert r2 is not runner and r2.stem.weight is qm.queue_network.feature_extractor.module.model.conv1.weight
    assert all(not p.requires_grad for p in qm.queue_network.parameters())
    assert len(qm.queue_network.vince_parameters()) == len(m.vince_parameters())


def test_split_dict_by_type_semantics():
    import vince_b200
    d = {"a": torch.arange(10), "b": torch.arange(20).view(10, 2)}
    out = vince_b200.VinceModel.split_dict_by_type(["x", "y"], [4, 6], d)
    assert [o["batch_type"] for o in out] == ["x", "y"]
    assert out[0]["a"].tolist() == [0, 1, 2, 3] and out[1]["b"].shape == (6, 2)
    one = vince_b200.VinceModel.split_dict_by_type(["images"], [10], d)
    assert one[0]["a"].shape == (10,) and "batch_types" not in one[0]


def test_stem_geometry_and_ring_slices():
    from vince_b200 import ops
    from vince_b200.distributed import ring_slices
    for H in (224, 225, 75, 64, 51):
        sg = ops.stem_geometry(H, H)
        P = (H + 6 - 7) // 2 + 1
        assert sg["P"] == P and sg["Ha"] == P + 3 and sg["Wb"] == sg["Q"] + 3
        g = sg["geom"]
        assert (g["H"] + g["pad_lo_h"] + g["pad_hi_h"] - g["R"]) // g["stride"] + 1 == P and g["pad_hi_h"] >= 0
        # overlapping-window layout: the last 64-element pixel of a row ends exactly at the row's last stored element
        assert (g["W"] - 1) * g["a_pixel_stride"] + g["Cin"] == g["a_row_stride"]
        assert g["a_img_stride"] == g["a_row_stride"] * g["H"]
    rng = np.random.RandomState(0)
    for _ in range(200):
        K = int(rng.randint(1, 40))
        q = vo.StorageQueue(K, 1, init=torch.zeros(K, 1))
        tail = 0
        for step in range(12):
            n = int(rng.randint(0, K + 1))
            items = torch.full((n, 1), float(step + 1))
            ref_before = q.vector_queue.clone()
            q.enqueue(items)
            slices, new_tail, wrapped = ring_slices(tail, n, K)
            mine = ref_before.clone()
            for src, dst, cnt in slices:
                mine[dst:dst + cnt] = items[src:src + cnt]
            assert torch.equal(mine, q.vector_queue)
            # the reference leaves tail == K (not 0) when a batch ends exactly at the end of the buffer
            assert new_tail == q.current_tail
            tail = new_tail
    with pytest.raises(ValueError):
        ring_slices(0, 5, 4)


def test_activation_arena_reuses_buffers():
    from vince_b200.encoder import _Arena
    a = _Arena("cpu")
    t1 = a.alloc((1000, 64), torch.float32)
    a.free(t1)
    t2 = a.alloc((2000, 64), torch.float16)        # same byte size: must reuse
    assert t2.data_ptr() == t1.data_ptr() and a.total == 1000 * 64 * 4
    t3 = a.alloc((10,), torch.float32)
    assert t3.data_ptr() != t2.data_ptr()


def test_two_pass_policy_picks_the_wide_small_k_expansions():
    """encoder.EncoderRunner._two_pass (host logic, no GPU): with the default policy the train-mode two-pass BatchNorm route
    (transposed statistics pass + apply epilogue) is taken by exactly the Bottleneck 1x1 expansions of the 56x56 and 28x28
    stages (K = 64 / 128, N = 4K: 7 of ResNet-50's 16), with VINCE_B200_TWOPASS=2 by all 16, never by a 3x3, a strided or a
    reducing convolution, and never in ResNet-18 (BasicBlocks have no 1x1 expansion)."""
    import vince_b200
    for backbone, expect1, expect2 in (("ResNet50", 7, 16), ("ResNet18", 0, 0)):
        args = make_args(backbone=backbone, batch_size=8, num_frames=2, queue_size=64, embedding_size=128, device="cpu")
        runner = vince_b200.VinceModel(args).feature_extractor.module.runner
        mains = [s for blk in runner.blocks for s in blk["convs"]]
        picked = {}
        for mode in (0, 1, 2):
            runner.two_pass = mode
            picked[mode] = [s for s in mains if runner._two_pass(s)]
        assert len(picked[0]) == 0 and len(picked[1]) == expect1 and len(picked[2]) == expect2
        for s in picked[2]:
            assert s.R == 1 and s.stride == 1 and s.Cout == 4 * s.K
        assert all(s.K <= 128 for s in picked[1])
        downs = [blk["down"] for blk in runner.blocks if blk["down"] is not None]
        runner.two_pass = 2
        assert len(downs) == (4 if backbone == "ResNet50" else 3)        # (downsample branches keep the raw route: encoder.py)


def test_graph_replay_runs_eagerly_when_disabled_or_profiled():
    """ops.GraphReplay (host logic, no GPU): with graphs disabled (VINCE_B200_GRAPH=0) or while bench.py collects
    per-launch events (ops.PROFILE) the launch list runs eagerly, `pre` first, in order, every call."""
    from vince_b200 import ops
    log = []
    rep = ops.GraphReplay([lambda: log.append("a"), lambda: log.append("b")], pre=lambda: log.append("pre"))
    saved_enabled, saved_profile = ops.GraphReplay.enabled, ops.PROFILE
    try:
        ops.GraphReplay.enabled = False
        rep()
        rep()
        assert log == ["pre", "a", "b"] * 2 and rep.graph is None
        ops.GraphReplay.enabled = True
        ops.PROFILE = []
        rep()
        assert log == ["pre", "a", "b"] * 3 and rep.graph is None and rep.calls == 0
    finally:
        ops.GraphReplay.enabled, ops.PROFILE = saved_enabled, saved_profile
    empty = ops.GraphReplay([])
    empty()
    assert empty.graph is None
