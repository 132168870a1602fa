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
