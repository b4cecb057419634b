"""The reference's own test problem (test/problems/sphere_regression.jl) through benlsip_b200: same call as BEnlsip.tralcnllss.
Needs a B200 (the library has no CPU path):  python examples/sphere_regression.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import benlsip_b200 as B

x_l, x_u = np.array([-2.0, -1.5, 0.0]), np.array([2.0, 1.5, 2.0])
A, b = np.array([[1.0, 2.0, -1.0]]), np.array([0.5])


def r(x):
    return np.array([x[0] ** 2 + x[1] ** 2 - 2 * x[0] + np.sin(x[0] + x[1]) - 1.5, x[0] * x[1] + 0.5 * np.cos(2 * x[0]) - 0.8,
                     (x[0] - 1.0) ** 2 + (x[1] - 0.5) ** 2 - x[2], x[2] ** 2 - x[0] + 0.3 * np.sin(x[2]) - 0.2])


def jac_r(x):
    return np.array([[2 * x[0] - 2 + np.cos(x[0] + x[1]), 2 * x[1] + np.cos(x[0] + x[1]), 0.0], [x[1] - np.sin(2 * x[0]), x[0], 0.0],
                     [2 * (x[0] - 1), 2 * (x[1] - 0.5), -1.0], [-1.0, 0.0, 2 * x[2] + 0.3 * np.cos(x[2])]])


def c(x):
    return np.array([x @ x - 3.0])


def jac_c(x):
    return (2.0 * x)[None, :]


if __name__ == "__main__":
    trace = {}
    x_sol, y_sol = B.tralcnllss(np.array([1.0, 0.5, 1.5]), r, jac_r, c, jac_c, A, b, x_l, x_u, max_outer_iter=100, max_inner_iter=250,
                                trace=trace)
    print("x =", x_sol, " y =", y_sol, " |c(x)| =", abs(c(x_sol)[0]), " outer iterations:", trace["outer_iters"])
