"""dm-sonnet 2.0.0 `moving_averages.ExponentialMovingAverage`, RESTATED (the package is not vendored by the reference
and not installable here): zero-debiased EMA,

    update(v): counter += 1; hidden -= (hidden - v) * (1 - decay); average = hidden / (1 - decay ** counter)
    __call__(v) = update(v); return average

`initialize(v)` creates zero `hidden` / `average` of v's shape once; `counter` is int64."""
import torch


class ExponentialMovingAverage:
    def __init__(self, decay, name=None):
        self._decay = decay
        self.name = name
        self._counter = torch.zeros((), dtype=torch.int64)
        self._hidden = None
        self.average = None

    def initialize(self, value):
        if self._hidden is None:
            self._hidden = torch.zeros_like(value)
            self.average = torch.zeros_like(value)

    def update(self, value):
        self.initialize(value)
        value = value.detach()
        self._counter += 1
        counter = self._counter.to(value.dtype)
        self._hidden = self._hidden - (self._hidden - value) * (1 - self._decay)
        self.average = self._hidden / (1. - torch.pow(torch.tensor(self._decay, dtype=value.dtype), counter))

    @property
    def value(self):
        return self.average.clone()

    def reset(self):
        self._counter.zero_()
        self._hidden.zero_()
        self.average.zero_()

    def __call__(self, value):
        self.update(value)
        return self.value
