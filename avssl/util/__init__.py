from .data_utils import get_keypadding_mask
from .model_utils import freeze_model, unfreeze_model
