"""Install the B200 drop-ins into an imported reference tree.

    import sys; sys.path.insert(0, "/path/to/end-to-end_asr_pytorch/src")
    import asr_b200.patch as patch
    patch.install()            # before the reference builds its model / solver
    # ... then run the reference's train.py main() unchanged

What gets replaced:
  transformer.cif_model.CIF_Model.cif        -> sm_100a CIF kernel pair
  transformer.loss.cal_ctc_ce_loss / cal_ctc_qua_ce_loss, ctcModel.loss.cal_loss
                                             -> fused CTC kernels
  transformer.attention.MultiheadAttention, ctcModel.attention.MultiHeadAttention
                                             -> tcgen05 attention core (same parameters / state_dict)
Every module that already did `from transformer.loss import ...` (the solvers, the
encoder / decoder) is re-pointed as well, so import order does not matter.
"""
import importlib
import sys


def _swap_everywhere(old, new):
    """Re-point every module-level name that is bound to `old`."""
    n = 0
    for mod in list(sys.modules.values()):
        d = getattr(mod, "__dict__", None)
        if not d:
            continue
        for k, v in list(d.items()):
            if v is old and v is not new:
                d[k] = new
                n += 1
    return n


def install(attention=True, verbose=False):
    here = __name__.rsplit(".", 1)[0]
    ours_loss = importlib.import_module(here + ".transformer.loss")
    ours_ctc_loss = importlib.import_module(here + ".ctcModel.loss")
    ours_cif = importlib.import_module(here + ".transformer.cif_model")
    done = {}

    ref_cif = importlib.import_module("transformer.cif_model")
    ref_cif.CIF_Model.cif = ours_cif.CIF_Model.cif
    ref_cif.CIF_Model.forward = ours_cif.CIF_Model.forward      # same ops, follows the input's device
    done["CIF_Model.cif"] = 1

    ref_loss = importlib.import_module("transformer.loss")
    for name in ("cal_ctc_ce_loss", "cal_ctc_qua_ce_loss"):
        done[name] = _swap_everywhere(getattr(ref_loss, name), getattr(ours_loss, name))
    try:
        ref_closs = importlib.import_module("ctcModel.loss")
        done["cal_loss"] = _swap_everywhere(ref_closs.cal_loss, ours_ctc_loss.cal_loss)
    except ImportError:
        pass

    if attention:
        ours_att = importlib.import_module(here + ".transformer.attention")
        ref_att = importlib.import_module("transformer.attention")
        done["MultiheadAttention"] = _swap_everywhere(ref_att.MultiheadAttention, ours_att.MultiheadAttention)
        try:
            ours_catt = importlib.import_module(here + ".ctcModel.attention")
            ref_catt = importlib.import_module("ctcModel.attention")
            done["MultiHeadAttention"] = _swap_everywhere(ref_catt.MultiHeadAttention, ours_catt.MultiHeadAttention)
        except ImportError:
            pass
    if verbose:
        print("asr_b200.patch:", done)
    return done
