"""The part of BIOINF_tesi/data_pipe that touches the hot path: the wire format of a batch.

The reference one-hot encodes every sequence with a freshly fitted sklearn OneHotEncoder per sample and epoch and moves
every sample to the device on its own (data_pipe/utils.py:268-276, dataprepare.py:398-412).  Here a data set is packed
ONCE -- features as fp32 [N, F], sequences as uint8 base codes [N, 256] (a, c, g, t = 0..3, the encoder's sorted category
order) -- kept resident on the device, and a batch is an index gather.  The pandas / imblearn / miceforest preprocessing
in front of it is out of scope."""
from .wire import encode_sequences, decode_onehot, BalancePos_BatchSampler, PackedDataset, DeviceLoader, build_loaders

__all__ = ['encode_sequences', 'decode_onehot', 'BalancePos_BatchSampler', 'PackedDataset', 'DeviceLoader', 'build_loaders']
