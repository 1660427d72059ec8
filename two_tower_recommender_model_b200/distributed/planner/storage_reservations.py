"""``HeuristicalStorageReservation(percentage=0.05)``
(/root/reference/03_model_training.py:807): fraction of device memory the planner
must leave free for activations and workspaces."""


class HeuristicalStorageReservation:
    def __init__(self, percentage: float, parameter_multiplier: float = 6.0, dense_tensor_estimate=None) -> None:
        assert 0.0 <= percentage <= 1.0
        self._percentage = percentage

    @property
    def percentage(self) -> float:
        return self._percentage


class FixedPercentageStorageReservation(HeuristicalStorageReservation):
    pass
