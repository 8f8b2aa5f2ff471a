"""``.bin`` / ``.idx`` indexed token datasets (Megatron format, byte compatible).

Parity: reference ``core/datasets/indexed_dataset.py`` -- header ``MMIDIDX\\x00\\x00``, ``<Q`` version
= 1, ``<B`` dtype code, ``<Q`` sequence count, ``<Q`` document count, then int32 sequence lengths,
int64 byte pointers, int64 document indices (:38,134-184,239-249); ``.bin`` = raw token array.
"""
from __future__ import annotations

import os
import struct
from typing import List, Optional, Tuple, Type

import numpy as np
import torch

_INDEX_HEADER = b"MMIDIDX\x00\x00"


class DType:
    codes = {1: np.uint8, 2: np.int8, 3: np.int16, 4: np.int32, 5: np.int64, 6: np.float64, 7: np.float32, 8: np.uint16}

    @classmethod
    def code_from_dtype(cls, dt) -> int:
        for c, d in cls.codes.items():
            if np.dtype(d) == np.dtype(dt):
                return c
        raise ValueError(f"unsupported dtype {dt}")

    @classmethod
    def dtype_from_code(cls, c: int):
        return cls.codes[c]

    @staticmethod
    def optimal_dtype(cardinality: Optional[int]):
        return np.uint16 if cardinality is not None and cardinality < 65500 else np.int32


def get_idx_path(prefix: str) -> str:
    return prefix + ".idx"


def get_bin_path(prefix: str) -> str:
    return prefix + ".bin"


class _IndexWriter:
    def __init__(self, idx_path: str, dtype):
        self.path, self.dtype = idx_path, dtype

    def write(self, sequence_lengths: List[int], sequence_modes, document_indices: List[int]):
        item = np.dtype(self.dtype).itemsize
        pointers = np.zeros(len(sequence_lengths), dtype=np.int64)
        if len(sequence_lengths) > 1:
            np.cumsum(np.array(sequence_lengths[:-1], dtype=np.int64) * item, out=pointers[1:])
        with open(self.path, "wb") as f:
            f.write(_INDEX_HEADER)
            f.write(struct.pack("<Q", 1))
            f.write(struct.pack("<B", DType.code_from_dtype(self.dtype)))
            f.write(struct.pack("<Q", len(sequence_lengths)))
            f.write(struct.pack("<Q", len(document_indices)))
            f.write(np.array(sequence_lengths, dtype=np.int32).tobytes(order="C"))
            f.write(pointers.tobytes(order="C"))
            f.write(np.array(document_indices, dtype=np.int64).tobytes(order="C"))
            if sequence_modes is not None:
                f.write(np.array(sequence_modes, dtype=np.int8).tobytes(order="C"))


class _IndexReader:
    def __init__(self, idx_path: str, multimodal: bool = False):
        with open(idx_path, "rb") as f:
            if f.read(9) != _INDEX_HEADER:
                raise ValueError(f"bad header in {idx_path}")
            (version,) = struct.unpack("<Q", f.read(8))
            if version != 1:
                raise ValueError(f"bad version {version} in {idx_path}")
            (code,) = struct.unpack("<B", f.read(1))
            self.dtype = DType.dtype_from_code(code)
            self.dtype_size = np.dtype(self.dtype).itemsize
            (self.sequence_count,) = struct.unpack("<Q", f.read(8))
            (self.document_count,) = struct.unpack("<Q", f.read(8))
            offset = f.tell()
        self._mmap = np.memmap(idx_path, mode="r", order="C")
        buf = memoryview(self._mmap)
        self.sequence_lengths = np.frombuffer(buf, dtype=np.int32, count=self.sequence_count, offset=offset)
        offset += self.sequence_lengths.nbytes
        self.sequence_pointers = np.frombuffer(buf, dtype=np.int64, count=self.sequence_count, offset=offset)
        offset += self.sequence_pointers.nbytes
        self.document_indices = np.frombuffer(buf, dtype=np.int64, count=self.document_count, offset=offset)
        offset += self.document_indices.nbytes
        self.sequence_modes = (np.frombuffer(buf, dtype=np.int8, count=self.sequence_count, offset=offset)
                               if multimodal else None)

    def __len__(self):
        return self.sequence_count


class IndexedDataset(torch.utils.data.Dataset):
    def __init__(self, path_prefix: str, multimodal: bool = False, mmap: bool = True):
        self.path_prefix, self.multimodal = path_prefix, multimodal
        self.index = _IndexReader(get_idx_path(path_prefix), multimodal)
        self._bin = np.memmap(get_bin_path(path_prefix), mode="r", order="C") if mmap else None
        self._mmap = mmap

    def __len__(self):
        return len(self.index)

    def get(self, idx: int, offset: int = 0, length: Optional[int] = None) -> np.ndarray:
        ptr, n = int(self.index.sequence_pointers[idx]), int(self.index.sequence_lengths[idx])
        if length is None:
            length = n - offset
        ptr += offset * self.index.dtype_size
        if self._mmap:
            return np.frombuffer(self._bin, dtype=self.index.dtype, count=length, offset=ptr)
        with open(get_bin_path(self.path_prefix), "rb") as f:
            f.seek(ptr)
            return np.frombuffer(f.read(length * self.index.dtype_size), dtype=self.index.dtype)

    def __getitem__(self, idx):
        if isinstance(idx, slice):
            return [self.get(i) for i in range(*idx.indices(len(self)))]
        return self.get(int(idx))

    @property
    def sequence_lengths(self):
        return self.index.sequence_lengths

    @property
    def document_indices(self):
        return self.index.document_indices

    @staticmethod
    def exists(path_prefix: str) -> bool:
        return os.path.exists(get_idx_path(path_prefix)) and os.path.exists(get_bin_path(path_prefix))


class IndexedDatasetBuilder:
    def __init__(self, bin_path: str, dtype=np.int32, multimodal: bool = False):
        self.f = open(bin_path, "wb")
        self.dtype, self.multimodal = dtype, multimodal
        self.sequence_lengths: List[int] = []
        self.document_indices: List[int] = [0]
        self.sequence_modes = [] if multimodal else None

    def add_item(self, tensor, mode: int = 0):
        arr = np.array(tensor.numpy() if torch.is_tensor(tensor) else tensor, dtype=self.dtype)
        self.f.write(arr.tobytes(order="C"))
        self.sequence_lengths.append(arr.size)
        if self.multimodal:
            self.sequence_modes.append(mode)

    def add_document(self, tensor, lengths: List[int], modes=None):
        arr = np.array(tensor.numpy() if torch.is_tensor(tensor) else tensor, dtype=self.dtype)
        self.f.write(arr.tobytes(order="C"))
        self.sequence_lengths.extend(lengths)
        self.document_indices.append(len(self.sequence_lengths))
        if self.multimodal:
            self.sequence_modes.extend(modes if modes is not None else [0] * len(lengths))

    def end_document(self):
        self.document_indices.append(len(self.sequence_lengths))

    def add_index(self, path_prefix: str):
        other = _IndexReader(get_idx_path(path_prefix), self.multimodal)
        if np.dtype(other.dtype) != np.dtype(self.dtype):
            raise ValueError("dtype mismatch when merging datasets")
        off = len(self.sequence_lengths)
        self.sequence_lengths.extend(other.sequence_lengths.tolist())
        self.document_indices.extend((off + other.document_indices)[1:].tolist())
        if self.multimodal:
            self.sequence_modes.extend(other.sequence_modes.tolist())
        with open(get_bin_path(path_prefix), "rb") as f:
            while True:
                chunk = f.read(1 << 24)
                if not chunk:
                    break
                self.f.write(chunk)

    def finalize(self, idx_path: str):
        self.f.close()
        _IndexWriter(idx_path, self.dtype).write(self.sequence_lengths, self.sequence_modes, self.document_indices)
