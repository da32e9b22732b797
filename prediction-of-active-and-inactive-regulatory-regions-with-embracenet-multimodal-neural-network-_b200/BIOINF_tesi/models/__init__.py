"""Mirror of BIOINF_tesi/models/__init__.py:1-15 for the classes on the EmbraceNet hot path."""
from .CNN_net import CNN
from .FF_net import FFNN
from .FFNN_NoTrain import FFNN_NoTrain
from .CNN_NoTrain import CNN_NoTrain
from .CNN_pre import CNN_pre, CNN_pre_NoTrain
from .FFNN_pre import FFNN_pre, FFNN_pre_NoTrain
from .EmbraceNetMultimodal import EmbraceNet, EmbraceNetMultimodal, EmbraceNetMultimodal_NoTrain
from .ConcatNetMultimodal import ConcatNetMultimodal, ConcatNetMultimodal_NoTrain

__all__ = ['CNN', 'FFNN', 'CNN_pre', 'FFNN_pre', 'EmbraceNetMultimodal', 'CNN_pre_NoTrain', 'FFNN_pre_NoTrain',
           'EmbraceNetMultimodal_NoTrain', 'FFNN_NoTrain', 'CNN_NoTrain', 'EmbraceNet', 'ConcatNetMultimodal', 'ConcatNetMultimodal_NoTrain']
