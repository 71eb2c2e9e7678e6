from .bags import BagShard, BagLoader, pack_bags  # noqa: F401
