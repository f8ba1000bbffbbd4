"""Drop-in replacements of models/general_cf/{lightgcn,simgcl,sgl,ncl,hccf}.py (same module and
class names, same constructor / forward / cal_loss / full_predict contracts)."""
