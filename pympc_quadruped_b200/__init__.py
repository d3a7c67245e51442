"""B200-native batched convex-MPC engine (drop-in for pympc-quadruped's linear_mpc/mpc.py path)."""
