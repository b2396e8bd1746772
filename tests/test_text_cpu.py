"""CPU suite of the text-prompt path: oracle restatement of CLIP.encode_text against the golden written from the
reference's own constructor (oracle/pin_against_reference.py text), and this repo's BPE tokenizer against the token
ids the reference's tokenizer produced for the same sentences (stored in the same golden file)."""
import pytest
import torch

from oracle import clipspm_oracle as O
from tests import helpers as H

TEXT_CLASSES = ["run", "jumping jacks", "pour water into a glass", "riding a bike"]  # pin_against_reference.py


def test_oracle_text_tower_matches_reference_golden():
    g = H.golden("text_tower_4cls")
    w = O.make_text_weights(512, seed=0)
    with torch.no_grad():
        mine = O.class_text_features(w, g["tokens"])
    assert H.rel_err(mine, g["text_features"]) < 1e-4


def test_oracle_text_tower_is_causal_and_reads_the_eot_row():
    """a token after <eot> cannot change the feature (causal mask + EOT-row gather, clip_fsar.py:778-784,803)"""
    g = H.golden("text_tower_4cls")
    w = O.make_text_weights(512, seed=0)
    tok = g["tokens"][0, :2].clone()
    tok2 = tok.clone()
    eot = tok.argmax(-1)
    for i in range(tok.shape[0]):
        tok2[i, eot[i] + 1:] = 7     # garbage after the end-of-text token (smaller id than <eot>)
    with torch.no_grad():
        a, b = O.encode_text(w, tok), O.encode_text(w, tok2)
    assert torch.allclose(a, b, atol=1e-6)


def _tokenizer():
    from clip_spm_b200.tokenizer import ClipTokenizer, find_vocab
    try:
        find_vocab()
    except FileNotFoundError:
        pytest.skip("CLIP BPE vocabulary file not available on this machine")
    return ClipTokenizer()


def test_tokenizer_reproduces_reference_token_ids():
    from clip_spm_b200.tokenizer import PROMPT_TEMPLATES
    tk = _tokenizer()
    g = H.golden("text_tower_4cls")
    tokens = torch.stack([tk.tokenize([t.format(c) for c in TEXT_CLASSES]) for t in PROMPT_TEMPLATES])
    assert torch.equal(tokens, g["tokens"].int())


def test_tokenizer_edge_cases():
    tk = _tokenizer()
    t = tk.tokenize(["", "  Hello,   WORLD!! it's 42  "])
    assert t.shape == (2, 77) and t.dtype == torch.int32
    assert t[0, 0] == tk.sot and t[0, 1] == tk.eot and int(t[0, 2:].abs().sum()) == 0   # empty text: <sot><eot>
    assert int(t[1].argmax()) == int((t[1] != 0).sum()) - 1                              # <eot> is the last non-pad id
    assert torch.equal(tk.tokenize("hello, world!! it's 42"), t[1:2])                     # case / whitespace folding
    with pytest.raises(RuntimeError):
        tk.tokenize("word " * 100)
    long = tk.tokenize("word " * 100, truncate=True)
    assert long[0, -1] == tk.eot
