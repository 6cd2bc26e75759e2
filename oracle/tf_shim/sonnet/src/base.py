class Module:
    """snt.Module: only the constructor signature is used by vq_layers.py."""

    def __init__(self, name=None):
        self.name = name
