"""setup.py name of the Efficient_TT extension (Efficient_TT/setup.py); same ops."""
from efficient_tt_table import *  # noqa: F401,F403
from efficient_tt_table import (Eff_TT_backward, Eff_TT_forward, Fused_Eff_TT_backward,  # noqa
                                Fused_Extra_Eff_TT_backward, init_cuda)
