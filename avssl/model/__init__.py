from .kwClip import *
