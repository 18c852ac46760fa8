"""Reference-side harnesses: the UNMODIFIED reference installed into baseline/_ref (scripts/install_reference.sh) and the
code that drives it (never imported by blobctrl_b200/)."""
