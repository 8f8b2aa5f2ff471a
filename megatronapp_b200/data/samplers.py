"""Batch samplers sharded over data-parallel ranks (reference ``legacy/data/data_samplers.py``:
``MegatronPretrainingSampler`` single-pass and ``MegatronPretrainingRandomSampler`` cyclic)."""
from __future__ import annotations

import torch


class MegatronPretrainingSampler:
    def __init__(self, total_samples, consumed_samples, micro_batch_size, data_parallel_rank,
                 data_parallel_size, drop_last=True):
        if total_samples <= 0:
            raise ValueError(f"no sample to consume: {total_samples}")
        if data_parallel_rank >= data_parallel_size:
            raise ValueError("data_parallel_rank must be smaller than data_parallel_size")
        self.total_samples, self.consumed_samples = total_samples, consumed_samples
        self.micro_batch_size = micro_batch_size
        self.data_parallel_rank, self.data_parallel_size = data_parallel_rank, data_parallel_size
        self.micro_batch_times_data_parallel_size = micro_batch_size * data_parallel_size
        self.drop_last = drop_last

    def __len__(self):
        return self.total_samples

    def get_start_end_idx(self):
        s = self.data_parallel_rank * self.micro_batch_size
        return s, s + self.micro_batch_size

    def __iter__(self):
        batch = []
        for idx in range(self.consumed_samples, self.total_samples):
            batch.append(idx)
            if len(batch) == self.micro_batch_times_data_parallel_size:
                s, e = self.get_start_end_idx()
                yield batch[s:e]
                batch = []
        if batch and not self.drop_last:
            s, e = self.get_start_end_idx()
            yield batch[s:e]


class MegatronPretrainingRandomSampler:
    def __init__(self, dataset, total_samples, consumed_samples, micro_batch_size, data_parallel_rank,
                 data_parallel_size, data_sharding=True, seed=1234):
        self.dataset, self.total_samples, self.consumed_samples = dataset, total_samples, consumed_samples
        self.micro_batch_size = micro_batch_size
        self.data_parallel_rank, self.data_parallel_size = data_parallel_rank, data_parallel_size
        self.data_sharding, self.seed = data_sharding, seed
        self.micro_batch_times_data_parallel_size = micro_batch_size * data_parallel_size
        self.last_batch_size = total_samples % self.micro_batch_times_data_parallel_size

    def __len__(self):
        return self.total_samples

    def __iter__(self):
        active = self.total_samples - self.last_batch_size
        self.epoch = self.consumed_samples // active
        cur = self.consumed_samples % active
        g = torch.Generator()
        g.manual_seed(self.seed + self.epoch)
        if self.data_sharding:
            bucket = (self.total_samples // self.micro_batch_times_data_parallel_size) * self.micro_batch_size
            off = cur // self.data_parallel_size
            start = self.data_parallel_rank * bucket
            idx_range = [start + x for x in torch.randperm(bucket, generator=g).tolist()[off:]]
        else:
            full = (self.total_samples // self.micro_batch_size) * self.micro_batch_size
            r = torch.randperm(full, generator=g).tolist()
            idx_range = r[cur:][self.data_parallel_rank::self.data_parallel_size]
        batch = []
        for idx in idx_range:
            batch.append(idx)
            if len(batch) == self.micro_batch_size:
                self.consumed_samples += self.micro_batch_times_data_parallel_size
                yield batch
                batch = []


def build_pretraining_data_loader(dataset, consumed_samples, micro_batch_size, dp_rank, dp_size,
                                  dataloader_type="single", num_workers=0, seed=1234, collate_fn=None):
    if dataset is None:
        return None
    if dataloader_type == "single":
        sampler = MegatronPretrainingSampler(len(dataset), consumed_samples, micro_batch_size, dp_rank, dp_size)
    elif dataloader_type == "cyclic":
        sampler = MegatronPretrainingRandomSampler(dataset, len(dataset), consumed_samples, micro_batch_size,
                                                   dp_rank, dp_size, seed=seed)
    elif dataloader_type == "external":
        return dataset
    else:
        raise ValueError(f"{dataloader_type} dataloader type is not supported")
    return torch.utils.data.DataLoader(dataset, batch_sampler=sampler, num_workers=num_workers,
                                       pin_memory=torch.cuda.is_available(), collate_fn=collate_fn,
                                       persistent_workers=num_workers > 0)
