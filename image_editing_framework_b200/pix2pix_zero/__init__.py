from .attention_control import MyAttnProcessor, prep_unet, restore_original_processors
from .sd_utils import P2P_Zero, P2P_Zero_XL, P2P_Zero_NTI, P2P_Zero_XL_NTI
