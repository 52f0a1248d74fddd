from ...models.layers import MLPLayers, activation_layer, kmeans, sinkhorn_algorithm  # noqa: F401
from .rq import ResidualVectorQuantizer  # noqa: F401
from .rqvae import RQVAE  # noqa: F401
from .vq import VectorQuantizer  # noqa: F401
