"""CPU oracle of the yahr render loop -- TEST INFRASTRUCTURE (see yahr_oracle.hpp)."""
