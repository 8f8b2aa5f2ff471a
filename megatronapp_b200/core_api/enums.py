"""Enumerations of the reference (``core/enums.py``, ``core/transformer/enums.py``, ``legacy/model/enums.py``).

``ModelType`` members compare and print as their names (``str(ModelType.encoder_and_decoder) == "encoder_and_decoder"``): the
schedules and ``get_model`` of this framework take either the enum or the plain string."""
import enum


class _NameEnum(str, enum.Enum):
    def __str__(self):
        return self.value

    def __format__(self, spec):
        return format(self.value, spec)


class ModelType(_NameEnum):
    encoder_or_decoder = "encoder_or_decoder"
    encoder_and_decoder = "encoder_and_decoder"
    retro_encoder = "retro_encoder"
    retro_decoder = "retro_decoder"


class Fp8Recipe(_NameEnum):
    delayed = "delayed"
    tensorwise = "tensorwise"
    mxfp8 = "mxfp8"
    blockwise = "blockwise"


class AttnType(enum.Enum):
    self_attn = 1
    cross_attn = 2


class AttnMaskType(enum.Enum):
    padding = 1
    causal = 2
    no_mask = 3
    padding_causal = 4
    arbitrary = 5


class AttnBackend(enum.Enum):
    flash = 1
    fused = 2
    unfused = 3
    local = 4
    auto = 5


class LayerType(enum.Enum):
    encoder = 1
    decoder = 2
    retro_encoder = 3
    retro_decoder = 4
    retro_decoder_with_retriever = 5


class PositionEmbeddingType(enum.Enum):
    rope = 1
    learned_absolute = 2
    none = 3
