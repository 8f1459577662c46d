from .register import register_attention_control, unregister_attention_control
from .attention_base import AttentionControl, EmptyControl, AttentionStore, AttentionControlEdit
from .attention_control import AttentionReplace, AttentionRefine, AttentionReweight
from .ptp_utils import LocalBlend, get_time_words_attention_alpha
from . import seq_aligner
from .sd_utils import P2P, P2P_NTI, P2P_XL, P2P_XL_NTI
