"""active_gym_b200 — B200-native batched implementation of Active-Gym's active-perception
observation path (frame-skip max-pool, grayscale + resize, frame-stack ring, foveal crop,
peripheral view, merge) behind the reference's gymnasium-style API.

Public names mirror ``active_gym/__init__.py:3-9,58-64`` of the reference.
"""
from .engine import ObservationPath, LUMA_RGB, LUMA_DMC  # noqa: F401

__all__ = ["ObservationPath", "LUMA_RGB", "LUMA_DMC"]
