"""Drop-in mirrors of `mantis.models.*` for the hot-path families (mllava, idefics2)."""
