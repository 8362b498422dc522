"""Drop-in mirrors of the reference's ``models`` package (same module and class names)."""
