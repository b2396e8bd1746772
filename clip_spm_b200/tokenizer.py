"""CLIP byte-level BPE tokenizer (host side of the text-prompt path).

Behavioural mirror of the reference's `SimpleTokenizer` / `tokenize` (models/clip_fsar.py:322-392, :144-180), written
from the published algorithm: lower-cased, whitespace-collapsed text is split by the CLIP regex, every piece is
mapped to printable byte symbols, its last symbol is tagged with `</w>`, and adjacent symbol pairs are merged in the
order of the merge table until no listed pair remains; ids follow the table
[256 byte symbols | the same with </w> | 48894 merges | <|startoftext|> | <|endoftext|>].

The merge table is the standard CLIP vocabulary file `bpe_simple_vocab_16e6.txt.gz` (1.3 MB).  It is NOT vendored
here: pass its path, set CLIPSPM_BPE_VOCAB, or keep it where the reference keeps it (models/ next to clip_fsar.py).
`ftfy.fix_text` (a unicode repair the reference applies first) is used when the module is installed and skipped
otherwise -- it is the identity on the plain-ASCII class names of the shipped configs."""
import gzip
import html
import os

import regex

SOT, EOT = "<|startoftext|>", "<|endoftext|>"
CONTEXT_LENGTH = 77
_N_MERGES = 49152 - 256 - 2          # clip_fsar.py:327
_PIECE = regex.compile(r"<\|startoftext\|>|<\|endoftext\|>|'s|'t|'re|'ve|'m|'ll|'d|[\p{L}]+|[\p{N}]|[^\s\p{L}\p{N}]+",
                       regex.IGNORECASE)


def _byte_symbols():
    """byte value -> printable unicode symbol: printable latin-1 bytes map to themselves, the 68 others to U+0100.."""
    keep = list(range(ord("!"), ord("~") + 1)) + list(range(0xA1, 0xAC + 1)) + list(range(0xAE, 0xFF + 1))
    table, extra = {}, 0
    for b in keep:
        table[b] = chr(b)
    for b in range(256):
        if b not in table:
            table[b] = chr(256 + extra)
            extra += 1
    # the id order of the vocabulary is: kept bytes in the order above, then the remapped ones in byte order
    order = keep + [b for b in range(256) if b not in set(keep)]
    return table, [table[b] for b in order]


def find_vocab(path=None):
    cands = [path, os.environ.get("CLIPSPM_BPE_VOCAB"),
             os.path.join(os.environ.get("CLIPSPM_REFERENCE", "/root/reference"), "models", "bpe_simple_vocab_16e6.txt.gz")]
    for c in cands:
        if c and os.path.exists(c):
            return c
    raise FileNotFoundError("CLIP BPE vocabulary (bpe_simple_vocab_16e6.txt.gz) not found: pass its path or set "
                            "CLIPSPM_BPE_VOCAB")


class ClipTokenizer:
    def __init__(self, vocab_path=None):
        self.byte_symbol, base = _byte_symbols()
        with gzip.open(find_vocab(vocab_path)) as f:
            lines = f.read().decode("utf-8").split("\n")
        merges = [tuple(l.split()) for l in lines[1:_N_MERGES + 1]]
        symbols = base + [s + "</w>" for s in base] + ["".join(m) for m in merges] + [SOT, EOT]
        self.ids = {s: i for i, s in enumerate(symbols)}
        self.rank = {m: i for i, m in enumerate(merges)}
        self.sot, self.eot = self.ids[SOT], self.ids[EOT]
        self._memo = {}

    def _merge(self, piece):
        """symbols of one regex piece after applying the merge table greedily by rank"""
        if piece in self._memo:
            return self._memo[piece]
        if piece in (SOT, EOT):
            return [piece]
        syms = list(piece[:-1]) + [piece[-1] + "</w>"]
        while len(syms) > 1:
            best, best_rank = None, None
            for a, b in zip(syms, syms[1:]):
                r = self.rank.get((a, b))
                if r is not None and (best_rank is None or r < best_rank):
                    best, best_rank = (a, b), r
            if best is None:
                break
            out, i = [], 0
            while i < len(syms):
                if i + 1 < len(syms) and syms[i] == best[0] and syms[i + 1] == best[1]:
                    out.append(best[0] + best[1])
                    i += 2
                else:
                    out.append(syms[i])
                    i += 1
            syms = out
        self._memo[piece] = syms
        return syms

    def encode(self, text):
        try:
            import ftfy
            text = ftfy.fix_text(text)
        except ImportError:
            pass
        text = html.unescape(html.unescape(text)).strip()
        text = regex.sub(r"\s+", " ", text).strip().lower()
        out = []
        for piece in _PIECE.findall(text):
            mapped = "".join(self.byte_symbol[b] for b in piece.encode("utf-8"))
            out.extend(self.ids[s] for s in self._merge(mapped))
        return out

    def tokenize(self, texts, context_length=CONTEXT_LENGTH, truncate=False):
        """-> int32 tensor [len(texts), context_length]: <sot> ids <eot> then zeros (clip_fsar.py:144-180)"""
        import torch
        if isinstance(texts, str):
            texts = [texts]
        res = torch.zeros(len(texts), context_length, dtype=torch.int32)
        for i, t in enumerate(texts):
            toks = [self.sot] + self.encode(t) + [self.eot]
            if len(toks) > context_length:
                if not truncate:
                    raise RuntimeError("Input %r is too long for context length %d" % (t, context_length))
                toks = toks[:context_length]
                toks[-1] = self.eot
            res[i, :len(toks)] = torch.tensor(toks, dtype=torch.int32)
        return res


# the 16 prompt templates of models/model_clipspm.py:45-49 (template -> one sentence per class name)
PROMPT_TEMPLATES = [
    "a photo of action {}", "a picture of action {}", "Human action of {}", "{}, an action", "{} this is an action",
    "{}, a video of action", "Playing action of {}", "{}", "Playing a kind of action, {}", "Doing a kind of action, {}",
    "Look, the human is {}", "Can you recognize the action of {}?", "Video classification of {}", "A video of {}",
    "The man is {}", "The woman is {}",
]
