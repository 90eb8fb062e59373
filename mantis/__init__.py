"""Alias shim: lets `from mantis.models.mllava import ...` / `from mantis.models.idefics2 import ...` (the import lines of
the reference's train drivers and examples) resolve to the B200-native implementations in `mantis_b200`."""
