from .register import regiter_attention_editor_diffusers, unregister_attention_control
from .attention_base import AttentionBase, AttentionStore
from .attention_control import (MutualSelfAttentionControl, MutualSelfAttentionControlUnion, MutualSelfAttentionControlMask,
                                MutualSelfAttentionControlMaskAuto)
