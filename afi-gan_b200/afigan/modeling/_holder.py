from ..functional import InferenceGraphs, PackedWeights


class Holder:
    """Non-module state attached to a Generator / discriminator stack (kept out of state_dict and of deepcopy's way)."""

    def __init__(self, n_rdb: int = 0):
        self.n_rdb = n_rdb
        self.packed = PackedWeights()
        self.graphs = InferenceGraphs()
        # Deferred weight gradients (opt-in, `Generator.deferred_weight_grads = True`): a neck calls the SAME interpolator many times per
        # step (28 in a BiFPN); instead of un-packing 23 gradient tensors per call and letting autograd add them up (28 x 23 tiny kernels),
        # every backward call adds into ONE packed accumulator and a callback queued on the autograd engine un-packs it into the parameters'
        # .grad once, when the backward pass ends.  Same sums; but the parameters' AccumulateGrad hooks do not fire, so keep it off under
        # torch DistributedDataParallel (FlatGradSync-style synchronisation after backward is fine).
        self.deferred = False
        self.acc = None
        self.pending = None

    def __deepcopy__(self, memo):
        h = Holder(self.n_rdb)
        h.deferred = self.deferred
        return h
