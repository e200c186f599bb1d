"""Drop-in for ``simulators_sc_ldpc/peeling_decoding/ber_sim.py`` (same argv, see ``main_simulate_sc_ldpc``)."""
from .peeling_decoding import main_simulate_sc_ldpc


def main():
    main_simulate_sc_ldpc()


if __name__ == '__main__':
    main()
