#!/usr/bin/env python
"""Quick start (needs one B200): the reference's smoke scenario (src/Pigeon.jl:34-57) for a batch of vehicles, then the same through the
callback entry point and a closed loop on the device.

    python pigeon.jl_b200/build.py        # once: nvcc -> pigeon.jl_b200/libpigeon_b200.so
    python examples/quickstart.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pigeon.jl_b200 as p

B = 4
X1 = p.X1()
traj = p.straight_trajectory(30.0, 5.0)                                  # straight 30 m at 5 m/s
mpc = p.BatchedCoupledTrajectoryTrackingMPC(X1, traj, B)                 # N_short = 10, N_long = 20, dt = 0.01 / 0.2 (reference defaults)
state = np.tile([0.0, 0.0, 0.0, 5.0, 0.0, 0.0], (B, 1))                  # (E, N, psi, Ux, Uy, r)
state[:, 0] = [0.0, 0.2, -0.2, 0.5]                                      # lateral offsets
control = np.zeros((B, 3))                                               # (delta, Fxf, Fxr)
mpc.set_state(state, control, other_car_state=np.tile([1e4, 1e4, 0.0, 5.0], (B, 1)))

# the reference's five calls (src/model_predictive_control.jl:70-78)
p.compute_time_steps(mpc, 0.0)
p.compute_linearization_nodes(mpc)
p.update_QP(mpc)
p.solve(mpc)
u = p.get_next_control(mpc)
print("get_next_control (delta, Fxf, Fxr):\n", u)
print("ADMM iterations:", mpc.stats()["iters"], "status:", mpc.stats()["status"])

# the ROS callback as one call: (delta, Fxf, Fxr, s_m, e_m) per vehicle (src/ros_integration.jl:48-151)
mpc.set_guards(nan_fallback=True, pause_below_speed=1.0)
print("from_autobox:\n", mpc.from_autobox(state, u, stamp=0.0))

# simulate (src/model_predictive_control.jl:80-100): 100 closed-loop steps on the device; returns the reference's (qs, xs, us, ps) histories,
# recorded on the device every 10th step
mpc.reset_solver(); mpc.reset_solved()
qs, xs, us, ps = p.simulate(mpc, state, control, 0.01, n_steps=100, stride=10)
q, u = mpc.get_state()
print("lateral error e of node 1 every 0.1 s (vehicle 3):", np.round(xs[:, 3, 5], 4))
print("after 1 s: lateral position E =", q[:, 0], " Fx =", u[:, 1] + u[:, 2], "(drag equilibrium 366.5 N)")

# pipelined host-buffer stepping: measured states go in, controls come out, three steps in flight (src/ros_integration.jl:50-53,96-99)
for k in range(6):
    mpc.step_submit(1.0 + 0.01 * k, q, u)
    if k >= 2:
        out = mpc.step_collect()
while True:
    try:
        out = mpc.step_collect()
    except p.PigeonError:
        break
print("last pipelined control:\n", out)
mpc.close()
