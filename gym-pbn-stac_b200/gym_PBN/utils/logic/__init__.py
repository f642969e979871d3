"""Boolean expression compiler for logic_func_data."""
