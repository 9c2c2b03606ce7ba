"""Builds the compiled parts of the oracle into oracle/_ref/ (git-ignored, travels with gpurun).
TEST INFRASTRUCTURE ONLY.  The reference itself is pure Python 2 + an absent third-party
package, so there is no reference source to compile here."""
from . import mapper_oracle


def build():
    return mapper_oracle.build_c()


if __name__ == '__main__':
    print(build())
