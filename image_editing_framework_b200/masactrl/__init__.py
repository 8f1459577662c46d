from .register import regiter_attention_editor_diffusers, unregister_attention_control
from .attention_base import AttentionBase, AttentionStore
from .attention_control import (MutualSelfAttentionControl, MutualSelfAttentionControlUnion, MutualSelfAttentionControlMask,
                                MutualSelfAttentionControlMaskAuto)
from .sd_utils import MasaCtrl, MasaCtrl_XL, MasaCtrl_NTI, MasaCtrl_XL_NTI
