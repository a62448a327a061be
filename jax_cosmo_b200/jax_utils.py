"""Parameter container mirroring jax_cosmo/jax_utils.py:4-30: positional args are the traced
parameters (`params`), keyword args the static configuration (`config`)."""


class container(object):
    def __init__(self, *args, **kwargs):
        self.params = args
        self.config = kwargs

    def __repr__(self):
        return str(self.params)

    def tree_flatten(self):
        return (self.params, self.config)

    @classmethod
    def tree_unflatten(cls, aux_data, children):
        return cls(*children, **aux_data)
