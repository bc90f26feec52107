import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import mcrg_b200
KC = float(-0.5 * np.log(1 + np.sqrt(2)))
with mcrg_b200.Context(4096, 40, seed=3) as ctx:
    ctx.set_couplings([KC]); ctx.init_hot(); ctx.sweep(30)
    ctx.set_update("cluster"); ctx.sweep(12); ctx.sync()
