"""Tokenised text classification dataset (reference src/dataset/AGNEWS.py:3-30): tokenise
per item, pad/truncate to ``max_length``, return input_ids / attention_mask / labels."""
from __future__ import annotations

import torch
from torch.utils.data import Dataset


class TokenizedTextDataset(Dataset):
    def __init__(self, texts, labels, tokenizer, max_length: int = 128):
        self.texts, self.labels, self.tok, self.max_length = list(texts), list(labels), tokenizer, max_length

    def __len__(self):
        return len(self.texts)

    def __getitem__(self, i):
        enc = self.tok(str(self.texts[i]), padding="max_length", truncation=True,
                       max_length=self.max_length, return_tensors="pt")
        return {"input_ids": enc["input_ids"].squeeze(0), "attention_mask": enc["attention_mask"].squeeze(0),
                "labels": torch.tensor(int(self.labels[i]), dtype=torch.long)}
