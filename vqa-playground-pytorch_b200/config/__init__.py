"""Config modules with the reference's surface (`--cf config.CoR2` style): module-level
hyper-parameters read by train.py plus `class Model` (reference: config/CoR2.py, config/ODA.py)."""
