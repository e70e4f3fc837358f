"""Drop-in ``embedding`` package: same module names as the reference's ``embedding/`` directory."""
