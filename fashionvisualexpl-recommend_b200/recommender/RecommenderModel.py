"""Base class of the recommender models: the attribute set of
src/recommender/RecommenderModel.py:5-25 without the Keras base class."""


class DeviceArray:
    """What ``predict_all()`` / ``call()`` return: a CUDA tensor with the ``.numpy()``
    accessor the reference's callers use (Evaluator.py:174,231; BPRMF.py:125)."""

    def __init__(self, tensor):
        self.tensor = tensor

    def numpy(self):
        return self.tensor.detach().cpu().numpy()

    @property
    def shape(self):
        return tuple(self.tensor.shape)

    def __array__(self, dtype=None):
        a = self.numpy()
        return a if dtype is None else a.astype(dtype)


class RecommenderModel:
    def __init__(self, data, params, *args, **kwargs):
        self.data = data
        self.num_items = data.num_items
        self.num_users = data.num_users
        self.params = params
        self.epochs = self.params.epochs
        self.batch_size = self.params.batch_size
        self.verbose = self.params.verbose
        self.restore_epochs = self.params.restore_epochs
        self.model_name = self.params.rec
        self.dataset_name = self.params.dataset

    def __call__(self, inputs, training=None, mask=None):
        return self.call(inputs, training=training, mask=mask)
