"""Reference arm support: the UNMODIFIED reference package (pure Python) copied to the git-ignored baseline/_ref/
so that it travels to the GPU box with the repo snapshot, and a timer for its own NumPy CPU path."""
