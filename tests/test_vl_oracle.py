"""Pins the Qwen2.5-VL oracle (the installed transformers, driven as the reference's wrapper drives it) to tests/golden/vl.pt and
checks the host-side bookkeeping of the native text encoder against the library (CPU; no kernels).  A transformers version whose
arithmetic or position handling differs from the one the goldens were made with fails here, not silently on the GPU."""
import pytest
import torch

from oracle import vl_oracle as VO


@pytest.fixture(scope="module")
def models():
    return VO.hf_model(torch.float32)


def rel(a, b):
    return ((a.float() - b.float()).norm() / b.float().norm()).item()


@pytest.mark.parametrize("case,with_image", [("image", True), ("text", False)])
def test_oracle_reproduces_goldens(golden, models, case, with_image):
    g = golden("vl")["cases"][case]
    inp = VO.inputs(with_image)
    assert inp["input_ids"].shape[1] == g["T"]
    h, last = VO.edit_forward(models, inp)
    assert torch.equal(h, last)                       # `hidden_states[-1]` IS the final-norm output the pipeline consumes
    assert rel(h[0], g["hidden_fp32"]) < 1e-5
    hseq, _ = VO.edit_forward(models, inp, mrope=False)
    assert rel(hseq[0], g["hidden_fp32_sequential"]) < 1e-5
    if with_image:                                    # 3-D vs sequential positions is a first-order effect, not a rounding detail
        assert rel(hseq[0], g["hidden_fp32"]) > 0.05
    else:
        assert rel(hseq[0], g["hidden_fp32"]) < 1e-5
    new, top2 = VO.generate(models, inp, 24)
    assert new.tolist() == g["tokens_fp32"].tolist()
    assert rel(g["hidden_bf16"], g["hidden_fp32"]) < 2e-2


def test_native_bookkeeping_matches_the_library(models):
    """Position ids (transformers 5.5 flavour), rotary tables, vision window index / rotary angles: restated in
    physicedit_b200/text_encoder.py, compared bit for bit with the library's own functions."""
    from physicedit_b200.text_encoder import mrope_position_ids, text_rope_tables, vision_rot_table, vision_window_index
    c = VO.native_config()
    for grid in ([(1, 12, 16)], [(1, 10, 12)], [(1, 8, 8), (1, 6, 20)]):
        n_img = [t * h * w // 4 for t, h, w in grid]
        ids = torch.cat([torch.tensor([5, 6, 7, VO.VISION_START])] + [torch.cat([torch.full((n,), VO.IMAGE_TOKEN), torch.tensor([VO.VISION_END, 9, VO.VISION_START])]) for n in n_img]
                        + [torch.tensor([11, 12, 13])]).view(1, -1)
        mm = (ids == VO.IMAGE_TOKEN).int()
        gt = torch.tensor(grid)
        pos_hf, delta_hf = models.model.get_rope_index(ids, mm_token_type_ids=mm, image_grid_thw=gt)
        pos, delta = mrope_position_ids(c, ids, grid, "mrope_hf55")
        assert torch.equal(pos_hf[:, 0], pos) and int(delta_hf) == delta
        # the 4.54 flavour differs only in the temporal index of the image tokens (start position, not 2 x start)
        p454, d454 = mrope_position_ids(c, ids, grid, "mrope")
        assert torch.equal(p454[1:], pos[1:]) and d454 == delta and torch.equal(p454[0][mm[0] == 0], pos[0][mm[0] == 0])
        first = int(mm[0].nonzero()[0])
        assert p454[0, first] == first
        win, cu = vision_window_index(c, grid)
        win_hf, cu_hf = models.model.visual.get_window_index(gt)
        assert torch.equal(win, win_hf) and cu.tolist() == torch.unique_consecutive(torch.tensor(cu_hf)).tolist()
        assert torch.equal(vision_rot_table(c, grid), models.model.visual.rot_pos_emb(gt))
        cos, sin = text_rope_tables(c, pos)
        chf, shf = models.model.language_model.rotary_emb(torch.zeros(1, 1), pos_hf)
        sec = [16, 24, 24] * 2
        assert torch.equal(torch.cat([m[i % 3] for i, m in enumerate(chf.split(sec, dim=-1))], dim=-1)[0], cos)
        assert torch.equal(torch.cat([m[i % 3] for i, m in enumerate(shf.split(sec, dim=-1))], dim=-1)[0], sin)
    seq, d0 = mrope_position_ids(c, torch.tensor([[4, 5, 6]]), None, "mrope")
    assert seq.tolist() == [[0, 1, 2]] * 3 and d0 == 0


def test_state_dict_layout_is_the_wrappers():
    """Same keys / shapes as transformers' Qwen2_5_VLForConditionalGeneration (what the wrapper subclasses), and the "diffusers"
    source layout converts into it (qwen_image_text_encoder_withdecode.py:283-297)."""
    from physicedit_b200.text_encoder import QwenImageTextEncoder, QwenImageTextEncoderStateDictConverter
    with torch.device("meta"):
        te = QwenImageTextEncoder(VO.native_config())
    hf = VO.hf_model(torch.float32)
    want = {k: tuple(v.shape) for k, v in hf.state_dict().items()}
    got = {k: tuple(v.shape) for k, v in te.state_dict().items()}
    assert got == want
    src = {}
    for k in want:                                     # the layout of the published checkpoint files
        if k.startswith("model.visual."):
            src[k[len("model."):]] = 0
        elif k.startswith("model.language_model."):
            src["model." + k[len("model.language_model."):]] = 0
        else:
            src[k] = 0
    assert set(QwenImageTextEncoderStateDictConverter().from_diffusers(src)) == set(want)
