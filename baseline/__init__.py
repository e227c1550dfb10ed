"""Reference arm support: `install_ref.install()` puts the UNMODIFIED wilson-labs/cola package into baseline/_ref
(git-ignored, but shipped to the GPU box with the repo snapshot like the built .so files)."""
