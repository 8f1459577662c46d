from .attention_control import MyAttnProcessor, prep_unet, restore_original_processors
