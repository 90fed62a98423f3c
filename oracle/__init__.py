"""CPU oracle — TEST INFRASTRUCTURE ONLY (see oracle/epi_oracle.c). Never imported by the product."""
