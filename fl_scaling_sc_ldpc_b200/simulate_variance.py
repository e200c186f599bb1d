"""Drop-in for ``simulators_sc_ldpc/peeling_decoding/simulate_variance.py`` (same argv, see ``main_simulate_variance``)."""
from .peeling_decoding import main_simulate_variance


def main():
    main_simulate_variance()


if __name__ == '__main__':
    main()
