from .dist import Dist
from .mlp import MLP
