from . import battle, pursuit, forest, double_attack
