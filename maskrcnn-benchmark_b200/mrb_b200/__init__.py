"""mrb_b200 -- build recipe and host-side harness around libmrb_b200.so (sm_100a)."""
