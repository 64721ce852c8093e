in0_liver.bed
