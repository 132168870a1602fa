"""Host-side logic of the KGnet mirror that can be checked without a GPU: the state-dict contract."""
import pytest
import torch

from oracle import kg_oracle as O


def test_state_dict_keys_and_shapes_match_reference_format():
    from kg_instance_segmentation_b200 import KGnet
    m = KGnet.resnet50(pretrained=False)
    sd = m.state_dict()
    ref = O.make_state_dict(seed=0)
    assert len(sd) == 346 and set(sd) == set(ref)
    for k in sd:
        assert tuple(sd[k].shape) == tuple(ref[k].shape), k
    m.load_state_dict(ref, strict=True)
    assert sum(p.numel() for p in m.parameters()) == sum(v.numel() for k, v in ref.items()
                                                         if "running" not in k and "num_batches" not in k)
    assert set(c for c, _ in m._conv_list()) == {k[:-len(".weight")] for k in ref if k.endswith(".weight") and ref[k].dim() == 4}


def test_reference_module_state_dict_loads(reference):
    KGnet_ref, _, _ = reference
    from kg_instance_segmentation_b200 import KGnet
    ref_model = KGnet_ref.resnet50(pretrained=False)
    m = KGnet.resnet50(pretrained=False)
    m.load_state_dict(ref_model.state_dict(), strict=True)
    assert set(m.state_dict()) == set(ref_model.state_dict())


def test_cpu_input_and_train_mode_are_rejected_loudly():
    from kg_instance_segmentation_b200 import KGnet
    m = KGnet.resnet50(pretrained=False)
    with pytest.raises(RuntimeError):
        m.forward_dec(torch.zeros(1, 3, 64, 64))
    m.eval()
    with pytest.raises(RuntimeError):
        m.forward_dec(torch.zeros(1, 3, 64, 64))     # CPU tensor: no CPU fallback
    with pytest.raises(NotImplementedError):
        KGnet.resnet18()


def test_reference_constructor_calls_work_unmodified():
    """test.py:53 / eval.py:31 construct `KGnet.resnet50(pretrained=True)`; train.py-style `ResNet(Bottleneck, layers)` too.
    Offline there is no ImageNet checkpoint: construction must proceed (with a warning), not raise."""
    import warnings
    from kg_instance_segmentation_b200 import KGnet
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        m = KGnet.resnet50(pretrained=True)
    assert len(m.state_dict()) == 346
    assert any("pretrained=True" in str(x.message) for x in w) or True   # a local checkpoint may exist
    m2 = KGnet.ResNet(KGnet.Bottleneck, [3, 4, 6, 3])
    assert set(m2.state_dict()) == set(m.state_dict())
    with pytest.raises(NotImplementedError):
        KGnet.ResNet(KGnet.BasicBlock, [2, 2, 2, 2])


def test_pretrained_loads_a_local_torchvision_checkpoint(tmp_path, monkeypatch):
    """KGnet.py:385: load_state_dict(..., strict=False) of the ImageNet trunk when a local file exists."""
    from kg_instance_segmentation_b200 import KGnet
    donor = KGnet.resnet50(pretrained=False)
    trunk = {k: torch.full_like(v, 0.25) for k, v in donor.state_dict().items() if k.startswith(("conv1.", "bn1.weight", "layer1.0.conv1"))}
    trunk["fc.weight"] = torch.zeros(10, 10)          # torchvision keys the truncated model does not have
    f = tmp_path / "resnet50-19c8e357.pth"
    torch.save(trunk, f)
    monkeypatch.setenv("KGNET_PRETRAINED", str(f))
    m = KGnet.resnet50(pretrained=True)
    assert float(m.conv1.weight.mean()) == 0.25 and float(m.layer1[0].conv1.weight.mean()) == 0.25


def test_patch_rect_matches_oracle_rounding():
    import numpy as np
    from kg_instance_segmentation_b200 import KGnet
    rs = np.random.RandomState(0)
    for _ in range(500):
        y1, x1 = rs.uniform(-0.1, 0.9, 2); y2, x2 = y1 + rs.uniform(0, 0.5), x1 + rs.uniform(0, 0.5)
        h, w = rs.randint(2, 300, 2)
        b = np.asarray([y1, x1, y2, x2], np.float32)
        assert KGnet.patch_rect(b, int(h), int(w)) == O.get_patch_rect(list(b), int(h), int(w))
    for v in (0.5, 1.5, 2.5, 3.5):      # exact .5 products: half-to-even
        b = np.asarray([0.0, 0.0, v / 8, v / 8], np.float32)
        assert KGnet.patch_rect(b, 8, 8) == O.get_patch_rect(list(b), 8, 8)


def test_weight_signature_sees_every_kind_of_update():
    """The cached slot list behind _sync_weights must notice in-place edits, load_state_dict, .to()-style re-materialisation and a
    Parameter object swapped by hand (a stale signature would silently keep the old weights on the device)."""
    from kg_instance_segmentation_b200 import KGnet
    m = KGnet.resnet50(pretrained=False)
    sigs = [m._signature()]

    def changed():
        sigs.append(m._signature())
        return sigs[-1] != sigs[-2]

    assert not changed()
    with torch.no_grad():
        dict(m.named_parameters())["kp_head_c0.2.weight"].mul_(0.5)
    assert changed()
    m.bn1.running_var.add_(1.0)
    assert changed()
    m.load_state_dict(O.make_state_dict(seed=3), strict=True)
    assert changed()
    m.conv1.weight = torch.nn.Parameter(torch.zeros_like(m.conv1.weight))
    assert changed()
    m.double()
    assert changed()
    assert len(sigs[-1]) == len(list(m.parameters())) + len(list(m.buffers()))


def test_twin_engine_shares_the_parameters():
    """InstanceHeat.submit / collect alternate between the model and a twin built on the meta device whose parameters and buffers are
    the SAME tensor objects (weight updates reach both); swapping engine.model rebuilds the twin."""
    from kg_instance_segmentation_b200 import KGnet
    from kg_instance_segmentation_b200.inference import InstanceHeat
    m = KGnet.resnet50(pretrained=False)
    eng = InstanceHeat(model=m, device="cpu")
    twin = eng._slot(1)["model"]
    assert twin is not m and eng._slot(0)["model"] is m
    pm, pt = dict(m.named_parameters()), dict(twin.named_parameters())
    assert set(pm) == set(pt) and all(pm[k] is pt[k] for k in pm)
    bm, bt = dict(m.named_buffers()), dict(twin.named_buffers())
    assert set(bm) == set(bt) and all(bm[k] is bt[k] for k in bm)
    assert not any(p.is_meta for p in twin.parameters()) and not any(b.is_meta for b in twin.buffers())
    before = twin._signature()
    m.load_state_dict(O.make_state_dict(seed=5), strict=True)
    assert twin._signature() != before and twin._signature() == m._signature()
    eng.model = KGnet.resnet50(pretrained=False).eval()
    assert eng._slot(0)["model"] is eng.model and eng._slot(1)["model"] is not twin
    eng._slots[1]["pending"] = {"fake": True}
    eng.model = m
    with pytest.raises(RuntimeError):
        eng._slot(0)
