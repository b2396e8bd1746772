"""GPU parity of the text-prompt tower (csrc/text_tower.cu) against the golden written from the reference's
constructor and against the oracle restatement of CLIP.encode_text."""
import pytest
import torch

from oracle import clipspm_oracle as O
from tests import helpers as H

pytestmark = pytest.mark.gpu

TOL = {"bf16": 5e-3, "fp32": 1e-4}   # "bf16" mode runs the text GEMMs in tf32 (fp32 activations throughout)


def _tower(precision, embed_dim=512):
    from clip_spm_b200 import TextTower
    return TextTower(O.make_text_weights(embed_dim, seed=0), precision=precision)


@pytest.mark.parametrize("precision", ["bf16", "fp32"])
def test_class_text_features_match_reference_golden(precision):
    g = H.golden("text_tower_4cls")
    tw = _tower(precision)
    out = tw.class_features_from_tokens(g["tokens"]).cpu()
    assert out.shape == g["text_features"].shape
    assert H.rel_err(out, g["text_features"]) < TOL[precision]


@pytest.mark.parametrize("precision", ["bf16", "fp32"])
def test_encode_text_matches_oracle(precision):
    g = H.golden("text_tower_4cls")
    tok = g["tokens"].reshape(-1, 77)[:9]
    w = O.make_text_weights(512, seed=0)
    with torch.no_grad():
        ref = O.encode_text(w, tok)
    tw = _tower(precision)
    assert H.rel_err(tw.encode_text(tok).cpu(), ref) < TOL[precision]
    one = tw.encode_text(tok[3:4]).cpu()                     # a single sentence
    assert H.rel_err(one, ref[3:4]) < TOL[precision]
    assert tw.encode_text(tok[:0]).shape == (0, 512)         # empty batch


def test_encode_text_rn50_width_and_chunking():
    """embed_dim 1024 (RN50 checkpoint) and more sentences than one device chunk (256)"""
    g = H.golden("text_tower_4cls")
    tok = g["tokens"].reshape(-1, 77)
    w = O.make_text_weights(1024, seed=0)
    with torch.no_grad():
        ref = O.encode_text(w, tok[:6])
    tw = _tower("bf16", 1024)
    many = tok.repeat(5, 1)[:300]
    out = tw.encode_text(many).cpu()
    assert out.shape == (300, 1024)
    assert H.rel_err(out[:6], ref) < TOL["bf16"]
    assert torch.allclose(out[256:300], out[:44], atol=1e-5)   # rows of the second chunk repeat rows of the first
    assert torch.allclose(out[64:128], out[:64], atol=1e-5)


def test_tokens_after_eot_do_not_matter():
    g = H.golden("text_tower_4cls")
    tok = g["tokens"][0, :3].clone()
    tok2 = tok.clone()
    eot = tok.argmax(-1)
    for i in range(tok.shape[0]):
        tok2[i, eot[i] + 1:] = 7
    tw = _tower("bf16")
    assert torch.allclose(tw.encode_text(tok), tw.encode_text(tok2), atol=1e-6)


def test_text_tower_errors_are_loud():
    from clip_spm_b200 import TextTower
    w = O.make_text_weights(512, seed=0)
    bad = dict(w)
    del bad["ln_final.weight"]
    with pytest.raises(RuntimeError, match="ln_final.weight"):
        TextTower(bad)
    tw = TextTower(w)
    with pytest.raises(RuntimeError):
        tw.encode_text(torch.zeros(2, 76, dtype=torch.int32))


CLASSES = ["run", "jumping jacks", "pour water into a glass", "riding a bike"]   # pin_against_reference.py


class GoldenTokenizer:
    """stands in for the BPE tokenizer where the vocabulary file is absent (the GPU box): returns the token ids the
    reference's tokenizer produced for exactly these sentences (stored in the golden file)"""

    def __init__(self, tokens):
        from clip_spm_b200.tokenizer import PROMPT_TEMPLATES
        self.table = {t.format(c): tokens[ti, ci] for ti, t in enumerate(PROMPT_TEMPLATES) for ci, c in enumerate(CLASSES)}

    def tokenize(self, texts):
        return torch.stack([self.table[t] for t in texts]).int()


def test_cnn_build_text_features_feeds_the_head():
    """class names -> text features -> episode head, end to end through the CNN mirror"""
    from clip_spm_b200.tokenizer import find_vocab
    g = H.golden("text_tower_4cls")
    try:
        find_vocab()
        tok = None                               # the real tokenizer where its vocabulary exists
    except FileNotFoundError:
        tok = GoldenTokenizer(g["tokens"])
    ci = H.case_inputs("head_5w5s_t8")
    m = H.build_cuda_model(ci, 1, "bf16")
    m.build_text_features(O.make_text_weights(512, seed=0), CLASSES, None, tokenizer=tok)
    assert H.rel_err(m.text_features_test.cpu(), g["text_features"]) < TOL["bf16"]
    # the head accepts the new class list: real labels now index 4 classes
    ep = ci["episode"]
    su, qu = ci["feats"]
    out = m.head(su.cuda()[None], qu.cuda()[None], ep["context_labels"].cuda()[None],
                 (ep["real_support_labels"] % 4).cuda()[None], (ep["real_target_labels"] % 4).cuda()[None])
    with torch.no_grad():
        st = O.head_forward(ci["weights"], g["text_features"], su, qu, ep["context_labels"],
                            ep["real_support_labels"] % 4, ep["real_target_labels"] % 4, O.DEFAULT_PARAMS, ci["single"])
    assert H.rel_err(out["logits"].cpu().reshape(st["logits"].shape), st["logits"]) < 2e-2
