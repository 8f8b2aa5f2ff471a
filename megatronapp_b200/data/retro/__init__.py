"""Retro retrieval preprocessing and dataset (reference ``core/datasets/retro/{config,db,index,query}`` + ``tools/retro``):

    chunk database  ->  ANN index (train / add)  ->  neighbour query of every pre-training sample  ->  RetroDataset

The reference leans on faiss (IVF-PQ on GPU) and HDF5; neither is part of this stack, so the index is implemented here
(`index.py`: k-means coarse quantiser + product quantiser + ADC search, all batched GEMMs / gathers that run on the B200
when one is present) and the artefacts are plain ``.npy`` memory maps."""
from .config import RetroPreprocessingConfig  # noqa: F401
from .db import ChunkDB, build_db  # noqa: F401
from .index import FlatIndex, IVFPQIndex, index_factory  # noqa: F401
from .query import RetroDataset, query_neighbors  # noqa: F401
