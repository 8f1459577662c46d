from .register import (load_source_latents_t, register_time, register_time_xl, register_attention_control_efficient, unregister_attention_control_efficient,
                       register_attention_control_efficient_xl, unregister_attention_control_efficient_xl,
                       register_conv_control_efficient, unregister_conv_control_efficient,
                       register_conv_control_efficient_xl, unregister_conv_control_efficient_xl)
from .sd_utils import PnP, PnP_XL, PnP_NTI, PnP_XL_NTI
