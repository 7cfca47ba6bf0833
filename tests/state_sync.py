"""Teacher-forcing helpers live in oracle/state_sync.py (the checker's side); re-exported for the tests."""
from oracle.state_sync import (ForegroundReconciler, SelectionReconciler, export_state_to_oracle,  # noqa: F401
                               load_state_from_oracle)
