from ..functional import InferenceGraphs, PackedWeights


class Holder:
    """Non-module state attached to a Generator / discriminator stack (kept out of state_dict and of deepcopy's way)."""

    def __init__(self, n_rdb: int = 0):
        self.n_rdb = n_rdb
        self.packed = PackedWeights()
        self.graphs = InferenceGraphs()

    def __deepcopy__(self, memo):
        return Holder(self.n_rdb)
