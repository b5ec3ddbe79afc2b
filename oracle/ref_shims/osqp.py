"""Stand-in module named `osqp` so the reference's `import osqp` resolves to the oracle's
RESTATEMENT of the OSQP algorithm (oracle/osqp_oracle.c).  This is NOT OSQP.  If a real osqp is
importable, oracle/ref_harness.py prefers it and records that."""
import numpy as np
from oracle import oracle as _orc

__version__ = "0+oracle-restatement"
DEFAULT_OVERRIDES = {}   # ref_harness sets e.g. eps_abs/eps_rel here for the parity setting
LOG = []                 # ref_harness can switch logging of every (P,q,A,l,u,x) on
LOG_ENABLED = False


class _Info:
    pass


class _Result:
    pass


class OSQP:
    def __init__(self):
        self._data = None
        self._settings = {}

    def setup(self, P=None, q=None, A=None, l=None, u=None, **settings):
        settings.pop("verbose", None)
        self._data = (P, np.asarray(q, float), A, np.asarray(l, float), np.asarray(u, float))
        self._settings = dict(DEFAULT_OVERRIDES)
        self._settings.update(settings)

    def solve(self):
        P, q, A, l, u = self._data
        n = P.shape[0]
        perm = None
        # the MPC QP (n = 5N+3, m = 8N+6) gets the stage-interleaved ordering for the band Cholesky
        if (n - 3) % 5 == 0 and A.shape[0] == 8 * ((n - 3) // 5) + 6:
            perm = _orc.stage_perm((n - 3) // 5)
        r = _orc.osqp_solve(P, q, A, l, u, perm=perm, **self._settings)
        res = _Result()
        res.info = _Info()
        res.info.status = r["status_str"]
        res.info.status_val = r["status"]
        res.info.iter = r["iter"]
        res.info.pri_res, res.info.dua_res, res.info.obj_val = r["pri_res"], r["dua_res"], r["obj"]
        res.info.rho_updates = r["rho_updates"]
        if np.isnan(r["x"][0]):
            # osqp 0.6 python wrapper returns arrays of None for infeasible problems
            res.x = np.array([None] * n)
            res.y = np.array([None] * A.shape[0])
        else:
            res.x, res.y = r["x"], r["y"]
        if LOG_ENABLED:
            LOG.append(dict(P=P.copy(), q=q.copy(), A=A.copy(), l=l.copy(), u=u.copy(), x=r["x"].copy(),
                            y=r["y"].copy(), status=r["status"], iter=r["iter"]))
        return res
