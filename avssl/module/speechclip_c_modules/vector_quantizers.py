from .my_vector_quantizer import *
