"""matplotlib stand-in: the reference imports pyplot at module top for debug plots
that the hot path never calls (test infrastructure only)."""
