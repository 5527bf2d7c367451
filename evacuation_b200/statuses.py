"""Pedestrian statuses and switch distances (src/env/env/statuses.py:16-27,
src/env/env/distances.py:17-21, src/env/constants.py:35-38)."""
from enum import Enum, auto


class UserEnum(Enum):
    @classmethod
    def all(cls):
        return list(cls)

    @classmethod
    def __len__(cls):
        return len(cls.all())


class Status(UserEnum):
    VISCEK = auto()    # 1: pedestrian under Vicsek rules
    FOLLOWER = auto()  # 2: follower of the leader particle (agent)
    EXITING = auto()   # 3: pedestrian in the exit zone
    ESCAPED = auto()   # 4: evacuated pedestrian


class SwitchDistances:
    to_leader: float = 0.2
    to_exit: float = 0.4
    to_escape: float = 0.01
    to_pedestrian: float = 0.1  # the "vision radius" of the Vicsek alignment
