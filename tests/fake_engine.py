"""TEST DOUBLE for MpcqEngine: same `solve` signature, answers from the oracle (reference
construction restated + exact solve).  Lets the host-side controller logic be tested on CPU.
Lives in tests/ only; the package has no such path."""
import numpy as np
import torch

from oracle import mpc_oracle as mo
from oracle.qp_exact import solve_qp_exact
from pympc_quadruped_b200.configs import extract_mpc_constants
from pympc_quadruped_b200.engine import SolveResult
from torch_statement import assemble_statement


class OracleEngine:
    def __init__(self, mpc_config, robot_config, dtype=torch.float32):
        self.c = extract_mpc_constants(mpc_config, robot_config)
        self.horizon = self.c["horizon"]
        self.dtype = dtype
        H = self.horizon
        self.Qbar = np.kron(np.identity(H), np.diag(self.c["q_diag"]))
        self.Rbar = np.kron(np.identity(H), np.diag(self.c["r_diag"]))
        self.calls = 0

    def assemble(self, *a, **k):
        assemble_statement(self.c, *a, **k)

    def solve(self, x0, r_feet, gait, x_ref, yaw=None, want=(), out=None):
        self.calls += 1
        B, H = x0.shape[0], self.horizon
        f = np.zeros((B, 12))
        u = np.zeros((B, 12 * H))
        for b in range(B):
            state = x0[b].cpu().numpy().astype(np.float32)
            y = float(yaw[b]) if yaw is not None else float(state[2])
            feet = r_feet[b].cpu().double().numpy().reshape(4, 3)
            Ac, Bc = mo.state_space_model(y, feet, self.c["inertia"], self.c["mass"])
            Ad, Bd = mo.discretize(Ac, Bc, self.c["dt"])
            Hm, g = mo.qp_cost(Ad, Bd, state, x_ref[b].cpu().numpy().astype(np.float32), self.Qbar, self.Rbar, H)
            _, _, ub = mo.qp_constraints(gait[b].cpu().numpy(), self.c["mu"], self.c["fz_max"], H)
            sol = solve_qp_exact(Hm, g, self.c["mu"], ub[4::5])
            assert sol.verified
            u[b] = sol.u
            f[b] = sol.u[:12]
        return SolveResult(forces=torch.as_tensor(f).to(self.dtype), u=torch.as_tensor(u).to(self.dtype), iters=None,
                           resid=None, status=None, active=None)
