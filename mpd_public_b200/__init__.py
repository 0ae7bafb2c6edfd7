"""Package marker (filled in below)."""
