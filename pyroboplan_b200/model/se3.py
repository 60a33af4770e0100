"""Minimal SE(3) value type (float64, host side).

Mirrors the small part of ``pinocchio.SE3`` that the reference's hot path touches
(``SE3(R, t)``, ``.rotation``, ``.translation``, ``*``, ``inverse``, ``act``, ``actInv``,
``Identity``) -- see /root/reference/src/pyroboplan/models/panda.py:98-154 (obstacle
placements) and /root/reference/src/pyroboplan/core/utils.py:253-255 (``data.oMf[...]``).
"""

import numpy as np


def rpy_to_matrix(roll, pitch, yaw):
    """URDF fixed-axis roll-pitch-yaw -> rotation matrix ``Rz(yaw) @ Ry(pitch) @ Rx(roll)``."""
    cr, sr = np.cos(roll), np.sin(roll)
    cp, sp = np.cos(pitch), np.sin(pitch)
    cy, sy = np.cos(yaw), np.sin(yaw)
    return np.array(
        [
            [cy * cp, cy * sp * sr - sy * cr, cy * sp * cr + sy * sr],
            [sy * cp, sy * sp * sr + cy * cr, sy * sp * cr - cy * sr],
            [-sp, cp * sr, cp * cr],
        ]
    )


def axis_angle_to_matrix(axis, angle):
    """Rodrigues rotation about a unit ``axis`` by ``angle``."""
    x, y, z = axis
    c, s = np.cos(angle), np.sin(angle)
    C = 1.0 - c
    return np.array(
        [
            [c + x * x * C, x * y * C - z * s, x * z * C + y * s],
            [y * x * C + z * s, c + y * y * C, y * z * C - x * s],
            [z * x * C - y * s, z * y * C + x * s, c + z * z * C],
        ]
    )


class SE3:
    """Rigid transform ``x -> R x + t``."""

    __slots__ = ("rotation", "translation")

    def __init__(self, rotation=None, translation=None):
        if rotation is None:
            rotation = np.eye(3)
        if translation is None:
            translation = np.zeros(3)
        rotation = np.asarray(rotation, dtype=np.float64)
        if rotation.shape == (4, 4) and translation is not None:
            translation = rotation[:3, 3]
            rotation = rotation[:3, :3]
        self.rotation = np.array(rotation, dtype=np.float64).reshape(3, 3)
        self.translation = np.array(translation, dtype=np.float64).reshape(3)

    @staticmethod
    def Identity():
        return SE3()

    @staticmethod
    def from_xyz_rpy(xyz, rpy):
        return SE3(rpy_to_matrix(*rpy), np.asarray(xyz, dtype=np.float64))

    def __mul__(self, other):
        if isinstance(other, SE3):
            return SE3(
                self.rotation @ other.rotation,
                self.rotation @ other.translation + self.translation,
            )
        return self.act(other)

    def inverse(self):
        rt = self.rotation.T
        return SE3(rt, -rt @ self.translation)

    def act(self, other):
        if isinstance(other, SE3):
            return self * other
        return self.rotation @ np.asarray(other, dtype=np.float64) + self.translation

    def actInv(self, other):
        if isinstance(other, SE3):
            return self.inverse() * other
        return self.rotation.T @ (np.asarray(other, dtype=np.float64) - self.translation)

    @property
    def homogeneous(self):
        m = np.eye(4)
        m[:3, :3] = self.rotation
        m[:3, 3] = self.translation
        return m

    def copy(self):
        return SE3(self.rotation.copy(), self.translation.copy())

    def isIdentity(self, tol=0.0):
        return bool(
            np.all(np.abs(self.rotation - np.eye(3)) <= tol)
            and np.all(np.abs(self.translation) <= tol)
        )

    def __repr__(self):
        return f"SE3(R=\n{self.rotation},\n t={self.translation})"
