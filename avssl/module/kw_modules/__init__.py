from . import TransformerModels
