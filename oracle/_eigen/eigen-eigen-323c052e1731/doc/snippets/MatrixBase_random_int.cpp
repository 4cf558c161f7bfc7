cout << VectorXi::Random(2) << endl;
