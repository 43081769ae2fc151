"""Data feed for the train step: the reference tokenises every caption again on every `__getitem__`, moves each
item to the device one by one and batches with `DataLoader(num_workers=0)` (CLIP-DDPM.py:167-197, 208-221) - at B200 step rates
(> 1 k captions/s) that host loop is the bottleneck. Here the captions are tokenised ONCE (by the caller, with whatever tokenizer
the reference would use) into `[N, MAX_LENGTH]` tensors that live on the device next to the precomputed CLIP features; a batch
is four gathers. The batch dict has the reference's keys, so `train_func(model, trainer, x)` / `validate` take it unchanged.
Works on CPU tensors too (host-logic tests)."""
from __future__ import annotations

from typing import Dict, Iterator, List, Optional, Sequence, Tuple

import torch


class DeviceCaptionDataset:
    """image_clip [N, 512] (one row per caption, as the reference's TensorDataset(image_set, text_set) holds), text_clip [N, 512],
    input_ids / attention_mask [N, MAX_LENGTH] int64, optional `image` names (list of N) for BLEU reference grouping."""

    def __init__(self, image_clip: torch.Tensor, text_clip: torch.Tensor, input_ids: torch.Tensor, attention_mask: torch.Tensor,
                 image: Optional[Sequence[str]] = None, text: Optional[Sequence[str]] = None, device=None):
        n = input_ids.shape[0]
        assert image_clip.shape[0] == text_clip.shape[0] == attention_mask.shape[0] == n
        assert input_ids.shape == attention_mask.shape
        dev = device if device is not None else input_ids.device
        self.image_clip = image_clip.to(dev, torch.float32).contiguous()
        self.text_clip = text_clip.to(dev, torch.float32).contiguous()
        self.input_ids = input_ids.to(dev, torch.int64).contiguous()
        self.attention_mask = attention_mask.to(dev, torch.int64).contiguous()
        self.image = list(image) if image is not None else None
        self.text = list(text) if text is not None else None
        self.device = self.input_ids.device

    @classmethod
    def from_captions(cls, captions: Sequence, images: Optional[Sequence[str]], image_clip: torch.Tensor, text_clip: torch.Tensor, tokenizer,
                      max_length: int = 16, device=None, chunk: int = 8192) -> "DeviceCaptionDataset":
        """Tokeniser adapter: what `FlickrCLIPDataset.__getitem__` (CLIP-DDPM.py:179-197) computes per item on every access, done ONCE for
        the whole caption list.

        * HF tokenizer (anything callable the way the reference calls `PreTrainedTokenizer`, :183):
          `tokenizer(text=..., return_tensors="pt", padding="max_length", truncation=True, max_length=MAX_LENGTH)` -> input_ids /
          attention_mask [N, MAX_LENGTH] ([CLS] ... [SEP] [PAD]...), batched `chunk` captions per call.
        * a vocabulary dict (the reference's `DictTokenizer` branch, :185-189; TRAIN_EMBEDDING): ids = [0] + [vocab.get(x, vocab['UNK']) for x in
          caption[:MAX_LENGTH-2]] + [1], padded with vocab['UNK'], mask 1 on the ids and 0 on the padding. `caption[:MAX_LENGTH-2]` is taken
          as the reference takes it: characters of a string, items of a list."""
        n = len(captions)
        vocab = tokenizer if isinstance(tokenizer, dict) else getattr(tokenizer, "dictionary", None)
        if isinstance(vocab, dict):
            unk = vocab["UNK"]
            ids = torch.full((n, max_length), unk, dtype=torch.int64)
            mask = torch.zeros((n, max_length), dtype=torch.int64)
            for i, cap in enumerate(captions):
                row = [0] + [vocab.get(x, unk) for x in cap[:max_length - 2]] + [1]
                ids[i, :len(row)] = torch.tensor(row, dtype=torch.int64)
                mask[i, :len(row)] = 1
        else:
            parts_i, parts_m = [], []
            for lo in range(0, n, chunk):
                tok = tokenizer(text=[str(c) for c in captions[lo:lo + chunk]], return_tensors="pt", padding="max_length", truncation=True,
                                max_length=max_length)
                parts_i.append(tok["input_ids"].to(torch.int64)); parts_m.append(tok["attention_mask"].to(torch.int64))
            ids, mask = torch.cat(parts_i), torch.cat(parts_m)
        return cls(image_clip, text_clip, ids, mask, image=images, text=[str(c) for c in captions], device=device)

    def __len__(self) -> int:
        return int(self.input_ids.shape[0])

    TENSOR_KEYS = ("image_clip", "text_clip", "input_ids", "attention_mask")

    def pin_memory(self) -> "DeviceCaptionDataset":
        """Host-resident variant (a caption set larger than HBM, or the bench's end-to-end leg): page-lock the four arrays so that
        batches gathered into pinned staging buffers (`loader(..., pin_staging=True)`) cross PCIe with asynchronous copies."""
        if self.device.type == "cpu" and torch.cuda.is_available():
            for k in self.TENSOR_KEYS:
                setattr(self, k, getattr(self, k).pin_memory())
        return self

    def batch(self, idx: torch.Tensor, out: Optional[Dict[str, torch.Tensor]] = None) -> Dict[str, object]:
        """Four gathers. `out` (host datasets): preallocated, pinned [B, ...] staging tensors the rows are gathered into."""
        idx = idx.to(self.device)
        if out is not None:
            for k in self.TENSOR_KEYS:
                torch.index_select(getattr(self, k), 0, idx, out=out[k])
            return dict(out)
        out: Dict[str, object] = {"image_clip": self.image_clip[idx], "text_clip": self.text_clip[idx], "input_ids": self.input_ids[idx],
                                  "attention_mask": self.attention_mask[idx]}
        if self.image is not None:
            out["image"] = [self.image[i] for i in idx.tolist()]
        if self.text is not None:
            out["text"] = [self.text[i] for i in idx.tolist()]
        return out

    def random_split(self, train_ratio: float, generator: Optional[torch.Generator] = None) -> Tuple["CaptionSubset", "CaptionSubset"]:
        """`train_len = int(len(dataset) * TRAIN_SET_RATIO)`; `random_split(dataset, [train_len, rest])` (CLIP-DDPM.py:212-213)."""
        n = len(self)
        perm = torch.randperm(n, generator=generator)
        n_train = int(n * train_ratio)
        return CaptionSubset(self, perm[:n_train]), CaptionSubset(self, perm[n_train:])


class CaptionSubset:
    def __init__(self, dataset: DeviceCaptionDataset, indices: torch.Tensor):
        self.dataset, self.indices = dataset, indices.clone()

    def __len__(self) -> int:
        return int(self.indices.numel())

    def loader(self, batch_size: int, shuffle: bool = False, generator: Optional[torch.Generator] = None, rank: int = 0,
               world: int = 1, pin_staging: bool = False) -> "CaptionLoader":
        return CaptionLoader(self, batch_size, shuffle, generator, rank, world, pin_staging)


class CaptionLoader:
    """`DataLoader(subset, shuffle=..., batch_size=BATCH_SIZE, drop_last=True)` (CLIP-DDPM.py:220-221); re-iterable (a new
    permutation per epoch when shuffling). Under data parallelism every rank draws the same permutation (same generator seed) and
    takes batches rank, rank + world, ... so the global batch is world * batch_size disjoint captions."""

    def __init__(self, subset: CaptionSubset, batch_size: int, shuffle: bool, generator: Optional[torch.Generator], rank: int, world: int,
                 pin_staging: bool = False):
        self.subset, self.batch_size, self.shuffle, self.generator, self.rank, self.world = subset, batch_size, shuffle, generator, rank, world
        # Host-resident dataset: batches are gathered into one of two alternating page-locked staging sets, so the caller's
        # `.to(device, non_blocking=True)` is a true asynchronous copy and the next gather does not overwrite a batch still in flight.
        self._staging = None
        ds = subset.dataset
        if pin_staging and ds.device.type == "cpu":
            pin = torch.cuda.is_available()
            self._staging = [{k: torch.empty((batch_size,) + tuple(getattr(ds, k).shape[1:]), dtype=getattr(ds, k).dtype, pin_memory=pin)
                              for k in ds.TENSOR_KEYS} for _ in range(2)]
        self._flip = 0

    def __len__(self) -> int:
        return (len(self.subset) // self.batch_size) // self.world

    def __iter__(self) -> Iterator[Dict[str, object]]:
        idx = self.subset.indices
        if self.shuffle:
            idx = idx[torch.randperm(idx.numel(), generator=self.generator)]
        n_batches = (idx.numel() // self.batch_size) // self.world * self.world   # drop_last, and the same count on every rank
        for b in range(self.rank, n_batches, self.world):
            rows = idx[b * self.batch_size:(b + 1) * self.batch_size]
            if self._staging is not None and self.subset.dataset.image is None and self.subset.dataset.text is None:
                self._flip ^= 1
                yield self.subset.dataset.batch(rows, out=self._staging[self._flip])
            else:
                yield self.subset.dataset.batch(rows)


def synthetic_dataset(n: int, max_length: int = 16, vocab: int = 30522, clip_dim: int = 512, seed: int = 0, device="cpu") -> DeviceCaptionDataset:
    """Synthetic stand-in of SURVEY 8(d): random ids with ragged lengths, unit-norm random CLIP features, 5 captions per image."""
    g = torch.Generator().manual_seed(seed)
    ids = torch.randint(0, vocab, (n, max_length), generator=g)
    lens = torch.randint(6, max_length + 1, (n,), generator=g)
    mask = (torch.arange(max_length)[None, :] < lens[:, None]).to(torch.int64)
    img = torch.nn.functional.normalize(torch.randn((n + 4) // 5, clip_dim, generator=g), dim=-1).repeat_interleave(5, 0)[:n]
    txt = torch.nn.functional.normalize(torch.randn(n, clip_dim, generator=g), dim=-1)
    return DeviceCaptionDataset(img, txt, ids, mask, image=[f"img{i // 5}.jpg" for i in range(n)], device=device)
